"""Kernel-level timing of the panel-chain pieces through the C ABI (N = 1, CUDA events, uncontended device):
  * sb200_potrf_tile_{d,s} on one nb x nb tile:   default | SB200_DIAG_MW=1|2 | SB200_TILE_FUSED=1|2 (+ DIAG_MW)
  * sb200_trsm_batched_{d,s}: the potrf panel solve (Right, Lower, Trans, NonUnit; batch tiles of nb x nb),
    the LU row solve (Left, Lower, NoTrans, Unit) and the small in-panel solves (na = 32, 64):
    default | SB200_DIAG_MW=1 | SB200_TRSM_FUSED bits
All switches used here are read per call, so one process measures every variant.
usage: python scratch/bench_tile.py [nb=512] [batch=32]"""
import ctypes, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from slate_b200._lib import lib, c_i64, c_int, c_dbl, c_flt, c_ptr

torch.cuda.set_device(0)
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 512
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 32
SWITCHES = ("SB200_DIAG_MW", "SB200_TILE_FUSED", "SB200_TRSM_FUSED")


def setenv(env):
    for k in SWITCHES:
        os.environ.pop(k, None)
    os.environ.update(env)


def timed(fn, reps=20):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e30
    for _ in range(3):
        e0.record()
        for _ in range(reps):
            fn()
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / reps)
    return best * 1e3          # us


def potrf_tile(t):
    tdt = torch.float64 if t == "d" else torch.float32
    rng = np.random.default_rng(1)
    G = rng.random((nb, nb)); S = G @ G.T + nb * np.eye(nb)
    A0 = torch.from_numpy(np.asfortranarray(S).T.copy()).to(tdt).cuda()      # column-major tile == the transpose's rows
    A = torch.empty_like(A0)
    info = torch.zeros(1, dtype=torch.int32, device="cuda")
    f = getattr(lib, f"sb200_potrf_tile_{t}")
    f.argtypes = [c_int, c_i64, c_ptr, c_i64, c_ptr, c_ptr, c_ptr]; f.restype = c_int
    st = torch.cuda.current_stream().cuda_stream

    def run():
        A.copy_(A0)
        assert f(ord("L"), nb, A.data_ptr(), nb, info.data_ptr(), None, st) == 0
    copy_us = timed(lambda: A.copy_(A0))
    for tag, env in (("default", {}), ("diag_mw", {"SB200_DIAG_MW": "1"}), ("diag_mw_rsqrt", {"SB200_DIAG_MW": "2"}),
                     ("tile_fused", {"SB200_TILE_FUSED": "1"}), ("tile_fused_rsqrt", {"SB200_TILE_FUSED": "2"})):
        setenv(env)
        try:
            us = timed(run) - copy_us
            ok = int(info.cpu()[0]) == 0
        except Exception as ex:  # noqa: BLE001
            us, ok = float("nan"), str(ex)
        print(json.dumps({"kernel": f"potrf_tile_{t}", "nb": nb, "variant": tag, "us": round(us, 1), "ok": ok}), flush=True)


def trsm(t, side, uplo, op, diag, m, n, bt, variants):
    tdt = torch.float64 if t == "d" else torch.float32
    sc = c_dbl if t == "d" else c_flt
    na = m if side == "L" else n
    rng = np.random.default_rng(2)
    T = torch.from_numpy((rng.random((na, na)) / na + np.eye(na) * 2).T.copy()).to(tdt).cuda()
    B0 = torch.from_numpy(rng.random((bt, n, m))).to(tdt).cuda()           # bt column-major m x n tiles
    B = torch.empty_like(B0)
    ptrs = torch.tensor([B.data_ptr() + i * m * n * B.element_size() for i in range(bt)], dtype=torch.int64, device="cuda")
    f = getattr(lib, f"sb200_trsm_batched_{t}")
    f.argtypes = [c_int] * 5 + [c_i64, c_i64, sc, c_ptr, c_i64, c_ptr, c_i64, c_i64, c_ptr, c_ptr]; f.restype = c_int
    st = torch.cuda.current_stream().cuda_stream

    def run():
        B.copy_(B0)
        assert f(ord("C"), ord(side), ord(uplo), ord(op), ord(diag), m, n, sc(1.0), T.data_ptr(), na, ptrs.data_ptr(), m, bt,
                 None, st) == 0
    copy_us = timed(lambda: B.copy_(B0))
    for tag, env in variants:
        setenv(env)
        try:
            us = timed(run) - copy_us
        except Exception as ex:  # noqa: BLE001
            us = float("nan"); tag += f" FAILED {ex}"
        print(json.dumps({"kernel": f"trsm_{t} {side}{uplo}{op}{diag} m={m} n={n} batch={bt}", "variant": tag, "us": round(us, 1)}),
              flush=True)


for t in ("d", "s"):
    potrf_tile(t)
    trsm(t, "R", "L", "T", "N", nb, nb, batch, (("default", {}), ("diag_mw", {"SB200_DIAG_MW": "1"}),
                                                ("fused", {"SB200_TRSM_FUSED": "1"}),
                                                ("fused+diag_mw", {"SB200_TRSM_FUSED": "1", "SB200_DIAG_MW": "1"})))
    trsm(t, "L", "L", "N", "U", nb, nb, 1, (("default", {}), ("diag_mw", {"SB200_DIAG_MW": "1"}),
                                            ("fused", {"SB200_TRSM_FUSED": "2"}),
                                            ("fused+diag_mw", {"SB200_TRSM_FUSED": "2", "SB200_DIAG_MW": "1"})))
    for na in (32, 64):
        trsm(t, "L", "L", "N", "U", na, na, 1, (("default", {}), ("diag_mw", {"SB200_DIAG_MW": "1"}),
                                                ("direct", {"SB200_TRSM_FUSED": "4"})))
setenv({})
