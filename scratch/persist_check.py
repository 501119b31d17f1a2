"""Opt-in persistent trailing GEMM (SB200_GEMM_PERSIST=<reserved slots>): bitwise comparison with the default kernel
on potrf / gemm / getrf at sizes whose trailing launches exceed the slot count, then timings with phase breakdown.
usage: python scratch/persist_check.py            (spawns itself once per mode)"""
import os, subprocess, sys, json
HERE = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, HERE)


def child(tag):
    import numpy as np, torch
    import slate_b200.host as sl
    torch.cuda.set_device(0)
    out = {}
    n, nb = 6144, 512
    H = sl.HermitianMatrix(n, nb).generate("rand_dominant", 42)
    assert sl.potrf(H) == 0
    np.save(f"/tmp/pc_{tag}_potrf.npy", np.tril(H.to_host()))
    A = sl.Matrix(n, n, nb).generate("rand", 1); B = sl.Matrix(n, n, nb).generate("rand", 2); C = sl.Matrix(n, n, nb).generate("rand", 3)
    sl.gemm(3.1, A, B, 2.7, C)
    np.save(f"/tmp/pc_{tag}_gemm.npy", C.to_host())
    G = sl.Matrix(n, n, nb).generate("rand", 42)
    piv, info = sl.getrf(G)
    np.save(f"/tmp/pc_{tag}_getrf.npy", G.to_host())
    del H, A, B, C, G
    for routine, mk, fl in (("potrf", lambda m: sl.HermitianMatrix(m, 512).generate("rand_dominant", 42), lambda m: m ** 3 / 3),
                            ("getrf", lambda m: sl.Matrix(m, m, 512).generate("rand", 42), lambda m: 2 * m ** 3 / 3)):
        m = 32768
        M0 = mk(m); M = mk(m)
        best = 1e30
        for r in range(3):
            M.copy_from(M0)
            os.environ["SB200_PHASES"] = "1" if r == 2 else "0"
            (sl.potrf(M) if routine == "potrf" else sl.getrf(M))
            best = min(best, M.last_driver_ms)
        out[routine] = {"ms": best, "tflops": fl(m) / best / 1e9, "panel_ms": M.last_panel_ms}
        del M, M0
    print("RESULT", tag, json.dumps(out), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1:
        child(sys.argv[1])
        sys.exit(0)
    import numpy as np
    for tag, env in (("base", {}), ("persist", {"SB200_GEMM_PERSIST": "16"})):
        e = dict(os.environ, **env)
        r = subprocess.run([sys.executable, __file__, tag], env=e, capture_output=True, text=True, timeout=200)
        print(r.stdout[-1500:]); print(r.stderr[-2500:])
    for what in ("potrf", "gemm", "getrf"):
        a, b = np.load(f"/tmp/pc_base_{what}.npy"), np.load(f"/tmp/pc_persist_{what}.npy")
        print(what, "bitwise identical" if np.array_equal(a, b) else f"DIFFER max {np.abs(a - b).max():.3e}", flush=True)
