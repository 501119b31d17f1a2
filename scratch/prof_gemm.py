import sys, torch
sys.path.insert(0, "/root/repo")
from slate_b200._lib import lib, check, c_i64, c_int, c_dbl, c_ptr
dev = torch.device("cuda:0")
lib.sb200_gemm_batched_d.argtypes = [c_int]*3 + [c_i64]*3 + [c_dbl, c_ptr, c_i64, c_ptr, c_i64, c_dbl, c_ptr, c_i64, c_i64, c_ptr]
st = torch.cuda.current_stream().cuda_stream
batch, nb = 600, 512
A = torch.rand(batch, nb, nb, dtype=torch.float64, device=dev)
B = torch.rand(batch, nb, nb, dtype=torch.float64, device=dev)
C = torch.rand(batch, nb, nb, dtype=torch.float64, device=dev)
ptrs = lambda ts: torch.tensor([t.data_ptr() for t in ts], dtype=torch.int64, device=dev)
pA, pB, pC = ptrs(A), ptrs(B), ptrs(C)
for opA, opB in (("N", "T"), ("N", "N")):
    for _ in range(2):
        check(lib.sb200_gemm_batched_d(ord('C'), ord(opA), ord(opB), nb, nb, nb, -1.0, pA.data_ptr(), nb, pB.data_ptr(), nb, 1.0, pC.data_ptr(), nb, batch, st))
torch.cuda.synchronize()
for _ in range(2):
    torch.baddbmm(C, A, B.transpose(1, 2), beta=1.0, alpha=-1.0)
torch.cuda.synchronize()
