"""ncu target: one dgetrf n=4096 nb=512 (8 panels) -- per-launch durations of the LU panel chain."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import slate_b200.host as sl
torch.cuda.set_device(0)
G = sl.Matrix(4096, 4096, 512).generate("rand", 42)
piv, info = sl.getrf(G)
assert info == 0
torch.cuda.synchronize()
