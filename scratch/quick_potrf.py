import sys, time, torch, numpy as np
sys.path.insert(0, "/root/repo")
import slate_b200.host as sl
from oracle import slate_oracle as o
torch.cuda.set_device(0)
for (n, nb) in [(512, 128), (1000, 128), (2048, 256), (1536, 512)]:
    A = sl.HermitianMatrix(n, nb)
    A.generate("rand_dominant", 42)
    Ah = A.to_host()
    G = o.generate("rand_dominant", n, n, 42)
    gen_ok = np.array_equal(np.tril(Ah), np.tril(G))
    info = sl.potrf(A)
    L = np.tril(A.to_host())
    Afull = np.tril(G) + np.tril(G, -1).T
    Lref = np.linalg.cholesky(Afull)
    err = np.abs(L - Lref).max() / np.abs(Lref).max()
    res = np.abs(L @ L.T - Afull).max() / (n * np.abs(Afull).max())
    print(f"potrf n={n} nb={nb}: gen_bitexact={gen_ok} info={info} max|L-Lref|/|L|={err:.2e} resid={res:.2e} {A.last_driver_ms:.2f} ms", flush=True)
# non positive definite: info
n, nb = 512, 128
H = o.generate("rand", n, n, 1); H = H + H.T + n * np.eye(n); H[300, 300] = -5.0
A = sl.HermitianMatrix(n, nb); A.from_host(np.asfortranarray(H))
print("non-PD info =", sl.potrf(A), "(expect 301)")
# perf
for (n, nb) in [(8192, 512), (16384, 512), (32768, 512), (32768, 256)]:
    A = sl.HermitianMatrix(n, nb)
    for rep in range(2):
        A.generate("rand_dominant", 42)
        t0 = time.time(); info = sl.potrf(A); t1 = time.time()
        fl = o.flops_potrf(n)
        print(f"potrf n={n} nb={nb}: info={info} dev {A.last_driver_ms:.1f} ms {fl/A.last_driver_ms/1e9:.2f} TFLOP/s | wall {1e3*(t1-t0):.1f} ms", flush=True)
    A.close()
