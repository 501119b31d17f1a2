"""ncu target: genorm (max, fro) over 512 tiles of 512x512 FP64."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from slate_b200._lib import lib, c_i64, c_int, c_ptr
torch.cuda.set_device(0)
nb, batch = 512, 512
te = nb * nb
a = torch.rand(batch * te, dtype=torch.float64, device="cuda")
pa = torch.tensor([a.data_ptr() + 8 * te * i for i in range(batch)], dtype=torch.int64, device="cuda")
vals = torch.zeros(batch * nb, dtype=torch.float64, device="cuda")
f = lib.sb200_genorm_batched_d
f.argtypes = [c_int, c_int, c_i64, c_i64, c_ptr, c_i64, c_ptr, c_i64, c_i64, c_ptr]; f.restype = c_int
st = torch.cuda.current_stream().cuda_stream
for norm, ldv in (("M", 1), ("O", nb), ("M", 1), ("F", 2)):
    assert f(ord(norm), ord("M"), nb, nb, pa.data_ptr(), nb, vals.data_ptr(), ldv, batch, st) == 0
torch.cuda.synchronize()
