"""Does reserving WHOLE SMs for the Cholesky chain kernels remove the 4.4x slow-down they suffer next to the trailing
update (r02b: potrf_tile_d 566 us idle, 2477 us contended)?  CUDA green contexts split the 148 SMs into a small
partition (chain stream) and the rest (trailing-update stream); the same kernels are timed
    idle / contended with stream priorities only (round-1 scheme) / contended with the chain on its own SM partition.
usage: python scratch/bench_greenctx.py [nb=512]"""
import ctypes, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from cuda.bindings import driver as cu
import slate_b200.host as sl  # noqa: F401  (loads the library)
from slate_b200._lib import lib, c_i64, c_int, c_dbl, c_ptr

torch.cuda.set_device(0)
torch.zeros(1, device="cuda")
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 512
te = nb * nb


def ck(r):
    if isinstance(r, tuple):
        err, rest = r[0], r[1:]
    else:
        err, rest = r, ()
    if int(err) != 0:
        raise RuntimeError(f"CUDA driver error {err}")
    return rest[0] if len(rest) == 1 else rest


def partitions(n_small, flags=0):
    """(stream on n_small SMs, stream on the rest, sm counts) via two green contexts"""
    dev = ck(cu.cuDeviceGet(0))
    res = ck(cu.cuDeviceGetDevResource(dev, cu.CUdevResourceType.CU_DEV_RESOURCE_TYPE_SM))
    out = cu.cuDevSmResourceSplitByCount(1, res, flags, n_small)
    if int(out[0]) != 0:
        raise RuntimeError(f"split failed: {out[0]}")
    groups, nb_groups, remaining = out[1], out[2], out[3]
    small = groups[0]
    d_small = ck(cu.cuDevResourceGenerateDesc([small], 1))
    d_big = ck(cu.cuDevResourceGenerateDesc([remaining], 1))
    g_small = ck(cu.cuGreenCtxCreate(d_small, dev, cu.CUgreenCtxCreate_flags.CU_GREEN_CTX_DEFAULT_STREAM))
    g_big = ck(cu.cuGreenCtxCreate(d_big, dev, cu.CUgreenCtxCreate_flags.CU_GREEN_CTX_DEFAULT_STREAM))
    s_small = ck(cu.cuGreenCtxStreamCreate(g_small, cu.CUstream_flags.CU_STREAM_NON_BLOCKING, -1))
    s_big = ck(cu.cuGreenCtxStreamCreate(g_big, cu.CUstream_flags.CU_STREAM_NON_BLOCKING, 0))
    return int(s_small), int(s_big), small.sm.smCount, remaining.sm.smCount


# background load: a batch of nb x nb x nb tile GEMMs ('N','T'), ~20 ms per launch
bt = 2400
Abg = torch.rand(64 * te, dtype=torch.float64, device="cuda")
Cbg = torch.rand(bt * te, dtype=torch.float64, device="cuda")
pa = torch.tensor([Abg.data_ptr() + 8 * te * (i % 64) for i in range(bt)], dtype=torch.int64, device="cuda")
pb = torch.tensor([Abg.data_ptr() + 8 * te * ((i * 7) % 64) for i in range(bt)], dtype=torch.int64, device="cuda")
pc = torch.tensor([Cbg.data_ptr() + 8 * te * i for i in range(bt)], dtype=torch.int64, device="cuda")
gemm = lib.sb200_gemm_batched_d
gemm.argtypes = [c_int, c_int, c_int, c_i64, c_i64, c_i64, c_dbl, c_ptr, c_i64, c_ptr, c_i64, c_dbl, c_ptr, c_i64, c_i64, c_ptr]
gemm.restype = c_int
pt = lib.sb200_potrf_tile_d
pt.argtypes = [c_int, c_i64, c_ptr, c_i64, c_ptr, c_ptr, c_ptr]; pt.restype = c_int
tr = lib.sb200_trsm_batched_d
tr.argtypes = [c_int] * 5 + [c_i64, c_i64, c_dbl, c_ptr, c_i64, c_ptr, c_i64, c_i64, c_ptr, c_ptr]; tr.restype = c_int

rng = np.random.default_rng(1)
G = rng.random((nb, nb)); S = G @ G.T + nb * np.eye(nb)
A0 = torch.from_numpy(S).cuda(); A = torch.empty_like(A0)
info = torch.zeros(1, dtype=torch.int32, device="cuda")
Lref = np.linalg.cholesky(S)
batch = 32
Tm = torch.from_numpy((rng.random((nb, nb)) / nb + np.eye(nb) * 2).T.copy()).cuda()
B0 = torch.rand(batch * te, dtype=torch.float64, device="cuda"); B = torch.empty_like(B0)
bptrs = torch.tensor([B.data_ptr() + 8 * te * i for i in range(batch)], dtype=torch.int64, device="cuda")


def background(stream, n_launch):
    for _ in range(n_launch):
        assert gemm(ord("C"), ord("N"), ord("T"), nb, nb, nb, -1.0, pa.data_ptr(), nb, pb.data_ptr(), nb, 1.0, pc.data_ptr(), nb, bt,
                    stream) == 0


def timed(fn, s_chain, s_bg, contended, reps=10):
    ext = torch.cuda.ExternalStream(s_chain)
    torch.cuda.synchronize()
    with torch.cuda.stream(ext):
        fn(s_chain)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if contended:
        background(s_bg, 6)
        time.sleep(0.005)
    with torch.cuda.stream(ext):
        e0.record(ext)
        for _ in range(reps):
            fn(s_chain)
        e1.record(ext)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


def potrf_tile(s):
    A.copy_(A0)
    assert pt(ord("L"), nb, A.data_ptr(), nb, info.data_ptr(), None, s) == 0


def trsm(s):
    B.copy_(B0)
    assert tr(ord("C"), ord("R"), ord("L"), ord("T"), ord("N"), nb, nb, 1.0, Tm.data_ptr(), nb, bptrs.data_ptr(), nb, batch, None, s) == 0


def bg_rate(s):
    torch.cuda.synchronize(); t0 = time.perf_counter(); background(s, 3); torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) * 1e3 / 3
    return round(2.0 * nb ** 3 * bt / ms / 1e9, 2)


configs = [("priorities_only", None, 0)]
for n_small in (8, 16):
    configs.append((f"green_{n_small}", n_small, 0))
configs.append(("green_4_ignore_cosched", 4, int(cu.CUdevSmResourceSplit_flags.CU_DEV_SM_RESOURCE_SPLIT_IGNORE_SM_COSCHEDULING)))

for name, n_small, flags in configs:
    try:
        if n_small is None:
            s_hi = torch.cuda.Stream(priority=-1); s_lo = torch.cuda.Stream(priority=0)
            s_chain, s_bg, sm_small, sm_big = s_hi.cuda_stream, s_lo.cuda_stream, 148, 148
        else:
            s_chain, s_bg, sm_small, sm_big = partitions(n_small, flags)
    except Exception as ex:            # noqa: BLE001
        print(json.dumps({"config": name, "error": str(ex)}), flush=True)
        continue
    rec = {"config": name, "sm_chain": sm_small, "sm_trailing": sm_big, "background_tflops": bg_rate(s_bg)}
    for variant, env in (("default", {}), ("tile_fused", {"SB200_TILE_FUSED": "1"})):
        for k, v in env.items():
            os.environ[k] = v
        for c in (False, True):
            rec[f"potrf_tile_{variant}_{'contended' if c else 'idle'}_us"] = round(timed(potrf_tile, s_chain, s_bg, c), 1)
        torch.cuda.synchronize()
        ok = bool(np.abs(np.tril(A.cpu().numpy()) - Lref).max() <= 64 * np.finfo(np.float64).eps * np.abs(Lref).max())
        rec[f"potrf_tile_{variant}_ok"] = ok
        for k in env:
            os.environ.pop(k)
    # the wide panel solve stays on the big partition at high priority in the proposed scheme: time it there
    if n_small is None:
        for c in (False, True):
            rec[f"trsm_b32_{'contended' if c else 'idle'}_us"] = round(timed(trsm, s_chain, s_bg, c), 1)
    else:
        dev = ck(cu.cuDeviceGet(0))
        for c in (False, True):
            rec[f"trsm_b32_on_chain_partition_{'contended' if c else 'idle'}_us"] = round(timed(trsm, s_chain, s_bg, c), 1)
    print(json.dumps(rec), flush=True)
