"""One-process timing probe (N=1): potrf and the getrf variants (SB200_PANEL=1|2, SB200_GETRF_DIST=0|1)
with per-phase device timers (SB200_PHASES=1 -> one JSON line per driver call on stderr).
usage: python scratch/perf_probe.py [n] [nb]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import slate_b200.host as sl
torch.cuda.set_device(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
nb = int(sys.argv[2]) if len(sys.argv) > 2 else 512
fl_potrf = n ** 3 / 3 + n ** 2 / 2 + n / 6
fl_getrf = 2 * n ** 3 / 3 - n * n / 2 + 5 * n / 6


def run(tag, A, A0, fn, flops, reps=3):
    best = 1e30
    for r in range(reps):
        A.copy_from(A0)
        os.environ["SB200_PHASES"] = "1" if r == reps - 1 else "0"
        if r == reps - 1:
            sys.stderr.write(f"## {tag}\n"); sys.stderr.flush()
        fn(A)
        best = min(best, A.last_driver_ms)
    print(f"{tag}: n={n} nb={nb} best {best:.2f} ms = {flops / best / 1e9:.2f} TFLOP/s  panel_ms={A.last_panel_ms:.1f}", flush=True)


H0 = sl.HermitianMatrix(n, nb).generate("rand_dominant", 42)
H = sl.HermitianMatrix(n, nb)
run("potrf", H, H0, lambda a: sl.potrf(a), fl_potrf)
H.close(); H0.close()
G0 = sl.Matrix(n, n, nb).generate("rand", 42)
G = sl.Matrix(n, n, nb)
for panel, dist in [("1", "0"), ("2", "0"), ("1", "1"), ("2", "1")]:
    os.environ["SB200_PANEL"] = panel
    os.environ["SB200_GETRF_DIST"] = dist
    try:
        run(f"getrf panel={panel} dist={dist}", G, G0, lambda a: sl.getrf(a), fl_getrf, reps=2)
    except Exception as ex:  # noqa: BLE001
        print(f"getrf panel={panel} dist={dist}: FAILED {ex}", flush=True)
