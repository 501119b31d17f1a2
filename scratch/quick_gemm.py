import ctypes, sys, time, torch
sys.path.insert(0, "/root/repo")
from slate_b200._lib import lib, check, c_i64, c_int, c_dbl, c_ptr
dev = torch.device("cuda:0")
print(torch.cuda.get_device_name(0))
lib.sb200_gemm_batched_d.argtypes = [c_int]*3 + [c_i64]*3 + [c_dbl, c_ptr, c_i64, c_ptr, c_i64, c_dbl, c_ptr, c_i64, c_i64, c_ptr]
lib.sb200_fp64_peak_probe.argtypes = [c_int, c_int, c_int, c_ptr, ctypes.POINTER(c_dbl), c_ptr]
st = torch.cuda.current_stream().cuda_stream

def ptrs(ts):
    return torch.tensor([t.data_ptr() for t in ts], dtype=torch.int64, device=dev)

def run(opA, opB, m, n, k, batch, layout='C', alpha=-1.0, beta=1.0, reps=0):
    # column-major storage emulated with torch: a col-major m x k is a (k, m) row-major tensor
    def mk(rows, cols):  # col-major rows x cols, ld = rows
        return torch.rand(batch, cols, rows, dtype=torch.float64, device=dev)
    if layout == 'C':
        A = mk(m, k) if opA == 'N' else mk(k, m)
        B = mk(k, n) if opB == 'N' else mk(n, k)
        C = mk(m, n)
        Am = A.transpose(1, 2) if opA == 'N' else A          # math view m x k
        Bm = B.transpose(1, 2) if opB == 'N' else B          # math view k x n
        Cm = C.transpose(1, 2)
        lda, ldb, ldc = A.shape[2], B.shape[2], m
    else:
        A = torch.rand(batch, *((m, k) if opA == 'N' else (k, m)), dtype=torch.float64, device=dev)
        B = torch.rand(batch, *((k, n) if opB == 'N' else (n, k)), dtype=torch.float64, device=dev)
        C = torch.rand(batch, m, n, dtype=torch.float64, device=dev)
        Am = A if opA == 'N' else A.transpose(1, 2)
        Bm = B if opB == 'N' else B.transpose(1, 2)
        Cm = C
        lda, ldb, ldc = A.shape[2], B.shape[2], n
    ref = alpha * torch.bmm(Am, Bm) + beta * Cm
    pA, pB, pC = ptrs(A), ptrs(B), ptrs(C)
    check(lib.sb200_gemm_batched_d(ord(layout), ord(opA), ord(opB), m, n, k, alpha, pA.data_ptr(), lda, pB.data_ptr(), ldb, beta, pC.data_ptr(), ldc, batch, st))
    torch.cuda.synchronize()
    err = (Cm - ref).abs().max().item() / max(1.0, ref.abs().max().item())
    msg = f"{layout} {opA}{opB} m={m} n={n} k={k} batch={batch} relerr={err:.2e}"
    if reps:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(2):
            lib.sb200_gemm_batched_d(ord(layout), ord(opA), ord(opB), m, n, k, alpha, pA.data_ptr(), lda, pB.data_ptr(), ldb, beta, pC.data_ptr(), ldc, batch, st)
        e0.record()
        for _ in range(reps):
            lib.sb200_gemm_batched_d(ord(layout), ord(opA), ord(opB), m, n, k, alpha, pA.data_ptr(), lda, pB.data_ptr(), ldb, beta, pC.data_ptr(), ldc, batch, st)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        msg += f"  {ms:.3f} ms  {2.0*m*n*k*batch/ms/1e9:.2f} TFLOP/s"
        # cuBLAS reference
        for _ in range(2): torch.bmm(Am, Bm)
        e0.record()
        for _ in range(reps): torch.bmm(Am, Bm)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        msg += f" | cuBLAS bmm {ms:.3f} ms {2.0*m*n*k*batch/ms/1e9:.2f} TFLOP/s"
    print(msg, flush=True)
    return err

# peak probes
scratch = torch.zeros(16, dtype=torch.float64, device=dev)
fl = c_dbl(0)
for kind, name in ((0, "DMMA.8x8x4"), (1, "DFMA")):
    for cps in (1, 2, 4):
        lib.sb200_fp64_peak_probe(kind, 2000, cps, scratch.data_ptr(), ctypes.byref(fl), st)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        lib.sb200_fp64_peak_probe(kind, 20000, cps, scratch.data_ptr(), ctypes.byref(fl), st)
        e1.record(); torch.cuda.synchronize()
        print(f"peak probe {name} ctas/sm={cps}: {fl.value/e0.elapsed_time(e1)/1e9:.2f} TFLOP/s", flush=True)

bad = 0
for layout in "CR":
    for opA in "NT":
        for opB in "NT":
            for (m, n, k, b) in [(80, 64, 16, 3), (128, 64, 16, 2), (256, 256, 256, 4), (130, 70, 36, 3), (77, 53, 19, 2), (512, 512, 512, 2)]:
                e = run(opA, opB, m, n, k, b, layout)
                bad += e > 1e-13
print("BAD =", bad)
import os
print("CFG", os.environ.get("SB200_GEMM_CFG"))
run("N", "T", 512, 512, 512, 1953, reps=5)
run('N', 'N', 512, 512, 512, 1953, reps=5)
run('T', 'N', 512, 512, 512, 1953, reps=5)
run('N', 'N', 512, 512, 512, 1953, layout='R', reps=5)
run('N', 'T', 256, 256, 256, 2048, reps=5)
big = torch.rand(8192, 8192, dtype=torch.float64, device=dev)
for _ in range(2): big @ big
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): big @ big
e1.record(); torch.cuda.synchronize()
print(f"cuBLAS dgemm 8192^3: {2*8192**3*5/e0.elapsed_time(e1)/1e9:.2f} TFLOP/s")
