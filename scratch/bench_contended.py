"""Chain kernels timed while a trailing-update-like batched DGEMM saturates the GPU on a LOW-priority stream (the
situation inside potrf / getrf), next to the same kernels on an idle GPU:
    potrf_tile_d (nb x nb), trsm_d RLTN (batch tiles), one LU panel (m x nb) via sl.getrf on a tall matrix.
usage: python scratch/bench_contended.py [nb=512]"""
import ctypes, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import slate_b200.host as sl
from slate_b200._lib import lib, c_i64, c_int, c_dbl, c_ptr

torch.cuda.set_device(0)
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 512
lo, hi = -1, 0
try:
    lo, hi = torch.cuda.Stream.priority_range()      # (least, greatest)
except Exception:
    lo, hi = 0, -1
s_hi = torch.cuda.Stream(priority=-1)
s_lo = torch.cuda.Stream(priority=0)

# background load: a batch of nb x nb x nb tile GEMMs ('N','T'), ~20 ms per launch
bt = 2400
te = nb * nb
Abg = torch.rand(64 * te, dtype=torch.float64, device="cuda")
Cbg = torch.rand(bt * te, dtype=torch.float64, device="cuda")
pa = torch.tensor([Abg.data_ptr() + 8 * te * (i % 64) for i in range(bt)], dtype=torch.int64, device="cuda")
pb = torch.tensor([Abg.data_ptr() + 8 * te * ((i * 7) % 64) for i in range(bt)], dtype=torch.int64, device="cuda")
pc = torch.tensor([Cbg.data_ptr() + 8 * te * i for i in range(bt)], dtype=torch.int64, device="cuda")
gemm = lib.sb200_gemm_batched_d
gemm.argtypes = [c_int, c_int, c_int, c_i64, c_i64, c_i64, c_dbl, c_ptr, c_i64, c_ptr, c_i64, c_dbl, c_ptr, c_i64, c_i64, c_ptr]
gemm.restype = c_int


def background(n_launch):
    for _ in range(n_launch):
        assert gemm(ord("C"), ord("N"), ord("T"), nb, nb, nb, -1.0, pa.data_ptr(), nb, pb.data_ptr(), nb, 1.0, pc.data_ptr(), nb, bt,
                    s_lo.cuda_stream) == 0


def timed(fn, contended, reps=10):
    """per-call us on the high-priority stream; contended: the background GEMM is running throughout"""
    torch.cuda.synchronize()
    with torch.cuda.stream(s_hi):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if contended:
        background(6)
        time.sleep(0.005)         # let the first launch fill the machine
    with torch.cuda.stream(s_hi):
        e0.record(s_hi)
        for _ in range(reps):
            fn()
        e1.record(s_hi)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


# background alone
torch.cuda.synchronize(); t0 = time.perf_counter(); background(3); torch.cuda.synchronize()
bg_ms = (time.perf_counter() - t0) * 1e3 / 3
print(json.dumps({"background_gemm_ms_per_launch": round(bg_ms, 2), "tflops": round(2.0 * nb ** 3 * bt / bg_ms / 1e9, 2)}), flush=True)

# potrf tile
rng = np.random.default_rng(1)
G = rng.random((nb, nb)); S = G @ G.T + nb * np.eye(nb)
A0 = torch.from_numpy(S).cuda(); A = torch.empty_like(A0)
info = torch.zeros(1, dtype=torch.int32, device="cuda")
pt = lib.sb200_potrf_tile_d
pt.argtypes = [c_int, c_i64, c_ptr, c_i64, c_ptr, c_ptr, c_ptr]; pt.restype = c_int


def potrf_tile():
    A.copy_(A0)
    assert pt(ord("L"), nb, A.data_ptr(), nb, info.data_ptr(), None, s_hi.cuda_stream) == 0


for c in (False, True):
    print(json.dumps({"kernel": "potrf_tile_d", "contended": c, "us": round(timed(potrf_tile, c), 1)}), flush=True)

# potrf panel solve
for batch in (1, 32):
    T = torch.from_numpy((rng.random((nb, nb)) / nb + np.eye(nb) * 2).T.copy()).cuda()
    B0 = torch.rand(batch * te, dtype=torch.float64, device="cuda"); B = torch.empty_like(B0)
    ptrs = torch.tensor([B.data_ptr() + 8 * te * i for i in range(batch)], dtype=torch.int64, device="cuda")
    tr = lib.sb200_trsm_batched_d
    tr.argtypes = [c_int] * 5 + [c_i64, c_i64, c_dbl, c_ptr, c_i64, c_ptr, c_i64, c_i64, c_ptr, c_ptr]; tr.restype = c_int

    def trsm():
        B.copy_(B0)
        assert tr(ord("C"), ord("R"), ord("L"), ord("T"), ord("N"), nb, nb, 1.0, T.data_ptr(), nb, ptrs.data_ptr(), nb, batch, None,
                  s_hi.cuda_stream) == 0
    for c in (False, True):
        print(json.dumps({"kernel": f"trsm_d RLTN batch={batch}", "contended": c, "us": round(timed(trsm, c), 1)}), flush=True)

# one LU panel: getrf of an m x nb matrix (only the panel kernels run); the driver uses its own streams, so the
# contended figure is taken by wall clock around the call
for m in (16384, 65536):
    P0 = sl.Matrix(m, nb, nb).generate("rand", 42); P = sl.Matrix(m, nb, nb)
    for c in (False, True):
        P.copy_from(P0); sl.getrf(P)
        ms = []
        for _ in range(3):
            P.copy_from(P0); torch.cuda.synchronize()
            if c:
                background(4); time.sleep(0.005)
            os.environ["SB200_PHASES"] = "1"
            piv, inf = sl.getrf(P)
            os.environ["SB200_PHASES"] = "0"
            ms.append(P.last_driver_ms)
            torch.cuda.synchronize()
        print(json.dumps({"kernel": f"LU panel {m} x {nb}", "contended": c, "ms": round(min(ms), 3),
                          "us_per_column": round(min(ms) * 1e3 / nb, 2)}), flush=True)
    P0.close(); P.close()
