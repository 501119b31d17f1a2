"""ncu target: the tcgen05 TF32x3 kernel inside a real posv_mixed call and the geadd tile kernel.
usage: python scratch/prof_mixed.py [n]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import slate_b200.host as sl
from slate_b200._lib import lib, c_i64, c_int, c_dbl, c_ptr
torch.cuda.set_device(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
nb = 512
A = sl.HermitianMatrix(n, nb).generate("rand_dominant", 42)
B = sl.Matrix(n, 10, nb).generate("rand", 43)
X = sl.Matrix(n, 10, nb)
print("posv_mixed", sl.posv_mixed(A, B, X)[:2], flush=True)
# geadd over 512 tiles of 512x512 (3 x 1 GiB of traffic)
batch = 512
te = nb * nb
a = torch.rand(batch * te, dtype=torch.float64, device="cuda"); b = torch.rand(batch * te, dtype=torch.float64, device="cuda")
pa = torch.tensor([a.data_ptr() + 8 * te * i for i in range(batch)], dtype=torch.int64, device="cuda")
pb = torch.tensor([b.data_ptr() + 8 * te * i for i in range(batch)], dtype=torch.int64, device="cuda")
f = lib.sb200_geadd_batched_d
f.argtypes = [c_i64, c_i64, c_dbl, c_ptr, c_i64, c_dbl, c_ptr, c_i64, c_i64, c_ptr]; f.restype = c_int
st = torch.cuda.current_stream().cuda_stream
for _ in range(3):
    assert f(nb, nb, 2.0, pa.data_ptr(), nb, 0.5, pb.data_ptr(), nb, batch, st) == 0
torch.cuda.synchronize()
