"""ncu target: one dpotrf n=2048 nb=512 (4 diagonal tiles) -- per-launch durations of the panel chain."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import slate_b200.host as sl
torch.cuda.set_device(0)
H = sl.HermitianMatrix(2048, 512).generate("rand_dominant", 42)
assert sl.potrf(H) == 0
torch.cuda.synchronize()
