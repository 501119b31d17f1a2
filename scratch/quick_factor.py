import ctypes, sys, torch, numpy as np
sys.path.insert(0, "/root/repo")
from slate_b200._lib import lib, check, c_i64, c_int, c_dbl, c_ptr
dev = torch.device("cuda:0")
st = torch.cuda.current_stream().cuda_stream
lib.sb200_trsm_batched_d.argtypes = [c_int]*5 + [c_i64, c_i64, c_dbl, c_ptr, c_i64, c_ptr, c_i64, c_i64, c_ptr, c_ptr]
lib.sb200_potrf_tile_d.argtypes = [c_int, c_i64, c_ptr, c_i64, c_ptr, c_ptr, c_ptr]
torch.manual_seed(1)
bad = 0
def colmajor(t):  # torch (rows, cols) math matrix -> column-major storage tensor (cols, rows)
    return t.transpose(-1, -2).contiguous()
for layout in "CR":
  for side in "LR":
    for uplo in "LU":
      for op in "NT":
        for diag in "NU":
          for (m, n, batch) in [(512, 512, 5), (200, 136, 3), (64, 300, 2), (77, 53, 2)]:
            na = m if side == "L" else n
            T = torch.rand(na, na, dtype=torch.float64, device=dev) + na * torch.eye(na, dtype=torch.float64, device=dev)
            Tm = torch.tril(T) if uplo == "L" else torch.triu(T)
            if diag == "U":
                Tm = Tm - torch.diag(torch.diag(Tm)) + torch.eye(na, dtype=torch.float64, device=dev)
                Tm = torch.tril(Tm, -1) / na + torch.triu(Tm, 1) / na + torch.eye(na, dtype=torch.float64, device=dev)
                T = torch.tril(T, -1) / na + torch.triu(T, 1) / na + torch.diag(torch.diag(T))
            B = torch.rand(batch, m, n, dtype=torch.float64, device=dev)
            alpha = 0.7
            opT = Tm if op == "N" else Tm.T
            upper = (uplo == "U") != (op == "T")
            ref = torch.linalg.solve_triangular(opT.expand(batch, na, na).contiguous(), alpha * B, upper=upper, left=(side == "L"))
            if layout == "C":
                Ts = colmajor(T); Bs = colmajor(B); ldb = m
            else:
                Ts = T.contiguous(); Bs = B.contiguous(); ldb = n
            pB = torch.tensor([Bs[i].data_ptr() for i in range(batch)], dtype=torch.int64, device=dev)
            check(lib.sb200_trsm_batched_d(ord(layout), ord(side), ord(uplo), ord(op), ord(diag), m, n, alpha, Ts.data_ptr(), na, pB.data_ptr(), ldb, batch, None, st))
            torch.cuda.synchronize()
            out = Bs.transpose(-1, -2) if layout == "C" else Bs
            err = ((out - ref).abs().max() / ref.abs().max()).item()
            ok = err < 1e-12
            bad += (not ok)
            if not ok or (m == 512 and layout == "C"):
                print(f"trsm {layout} {side}{uplo}{op}{diag} m={m} n={n} batch={batch} relerr={err:.2e} {'OK' if ok else 'BAD'}", flush=True)
print("TRSM BAD =", bad)
for n in (512, 256, 200, 64, 37):
    G = torch.rand(n, n, dtype=torch.float64, device=dev)
    A = G @ G.T + n * torch.eye(n, dtype=torch.float64, device=dev)
    As = colmajor(A)  # symmetric anyway
    As_before = As.clone()
    info = torch.zeros(1, dtype=torch.int32, device=dev)
    check(lib.sb200_potrf_tile_d(ord("L"), n, As.data_ptr(), n, info.data_ptr(), None, st))
    torch.cuda.synchronize()
    L = torch.tril(As.T)
    ref = torch.linalg.cholesky(A)
    err = ((L - ref).abs().max() / ref.abs().max()).item()
    upper_untouched = torch.equal(torch.triu(As.T, 1), torch.triu(As_before.T, 1))
    print(f"potrf_tile n={n} relerr={err:.2e} info={info.item()} upper_untouched={upper_untouched}")
A = torch.eye(128, dtype=torch.float64, device=dev); A[70, 70] = -1.0
As = A.clone(); info = torch.zeros(1, dtype=torch.int32, device=dev)
check(lib.sb200_potrf_tile_d(ord("L"), 128, As.data_ptr(), 128, info.data_ptr(), None, st)); torch.cuda.synchronize()
print("potrf_tile non-PD info =", info.item(), "(expect 71)")
n = 512
G = torch.rand(n, n, dtype=torch.float64, device=dev); A = G @ G.T + n * torch.eye(n, dtype=torch.float64, device=dev)
info = torch.zeros(1, dtype=torch.int32, device=dev)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
Acp = [A.clone() for _ in range(12)]
for i in range(2): lib.sb200_potrf_tile_d(ord("L"), n, Acp[i].data_ptr(), n, info.data_ptr(), None, st)
e0.record()
for i in range(2, 12): lib.sb200_potrf_tile_d(ord("L"), n, Acp[i].data_ptr(), n, info.data_ptr(), None, st)
e1.record(); torch.cuda.synchronize()
print(f"potrf_tile 512: {e0.elapsed_time(e1)/10*1000:.1f} us")
batch = 63
B = torch.rand(batch, n, n, dtype=torch.float64, device=dev)
pB = torch.tensor([B[i].data_ptr() for i in range(batch)], dtype=torch.int64, device=dev)
L = torch.linalg.cholesky(A).T.contiguous()
for i in range(2): lib.sb200_trsm_batched_d(ord("C"), ord("R"), ord("L"), ord("T"), ord("N"), n, n, 1.0, L.data_ptr(), n, pB.data_ptr(), n, batch, None, st)
e0.record()
for i in range(10): lib.sb200_trsm_batched_d(ord("C"), ord("R"), ord("L"), ord("T"), ord("N"), n, n, 1.0, L.data_ptr(), n, pB.data_ptr(), n, batch, None, st)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)/10
print(f"trsm R/L/T 63 tiles of 512: {ms*1000:.1f} us  ({batch*n*n*n/ms/1e9:.2f} TFLOP/s)")
