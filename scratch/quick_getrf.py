import sys, time, torch, numpy as np
sys.path.insert(0, "/root/repo")
import slate_b200.host as sl
from oracle import slate_oracle as o
torch.cuda.set_device(0)
g = np.load("/root/repo/tests/golden/getrf_d.npz")
for (n, nb) in [(384, 128), (300, 128), (512, 512), (1024, 256), (2048, 512), (4096, 512)]:
    A = sl.Matrix(n, n, nb).generate("rand", 42)
    piv, info = sl.getrf(A)
    LU = A.to_host()
    A0 = o.generate("rand", n, n, 42)
    perm = o.pivots_to_perm(piv, n, nb)
    L = np.tril(LU, -1) + np.eye(n); U = np.triu(LU)
    res = np.abs(A0[perm] - L @ U).max() / (n * np.abs(A0).max())
    msg = f"getrf n={n} nb={nb}: info={info} |PA-LU|/(n|A|)={res:.2e} {A.last_driver_ms:.2f} ms"
    if n <= 2048:
        LUo, pivo, infoo = o.getrf(A0, nb, 32)
        same = [tuple(x) for c in piv for x in c] == [tuple(x) for c in pivo for x in c]
        msg += f" pivots==oracle(ib32):{same} |LU-LUo|/|LU|={np.abs(LU-LUo).max()/np.abs(LUo).max():.2e}"
    if (n, nb) == (384, 128):
        flat = np.array([x for c in piv for x in c])
        msg += f" pivots==reference golden:{np.array_equal(flat, g['piv'])} |LU-ref|={np.abs(LU-g['out']).max()/np.abs(g['out']).max():.2e}"
    print(msg, flush=True)
# zero column -> info
n, nb = 256, 64
A0 = o.generate("rand", n, n, 3); A0[:, 100] = 0.0
A = sl.Matrix(n, n, nb); A.from_host(np.asfortranarray(A0))
piv, info = sl.getrf(A)
print("zero-column info =", info, "(expect 101)")
for (n, nb) in [(8192, 512), (16384, 512), (32768, 512)]:
    A = sl.Matrix(n, n, nb)
    for rep in range(2):
        A.generate("rand", 42)
        t0 = time.time(); piv, info = sl.getrf(A); t1 = time.time()
        print(f"getrf n={n} nb={nb}: info={info} dev {A.last_driver_ms:.1f} ms {o.flops_getrf(n,n)/A.last_driver_ms/1e9:.2f} TFLOP/s | wall {1e3*(t1-t0):.1f} ms", flush=True)
    A.close()
