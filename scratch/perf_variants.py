"""One-process timing of the round-2 candidates against the default path (N = 1), per-phase device timers on stderr.
Every variant is a set of environment switches read by the library PER CALL or at first use; variants whose switch is
latched at first use are run in a FRESH process each:
    python scratch/perf_variants.py potrf 32768 512            # all potrf variants, one subprocess each
    python scratch/perf_variants.py getrf 32768 512
    python scratch/perf_variants.py one potrf 32768 512        # (internal) run with the current environment
"""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

POTRF = [("default", {}),
         ("diag_division_form", {"SB200_DIAG_RSQRT": "0"}),
         ("tile_fused", {"SB200_TILE_FUSED": "1"}),
         ("tile_fused_rsqrt", {"SB200_TILE_FUSED": "2"}),
         ("no_fused_solves", {"SB200_TRSM_FUSED": "0"}),
         ("chain_on_4_sms", {"SB200_CHAIN_SMS": "4"}),
         ("chain_on_8_sms", {"SB200_CHAIN_SMS": "8"})]
GETRF = [("default", {}),
         ("no_folded_updates", {"SB200_PANEL_FUSE": "0"}),
         ("rows_in_shared_memory", {"SB200_PANEL_V4": "0"}),
         ("cooperative_barrier_kernel", {"SB200_PANEL_V3": "0"}),
         ("untransposed_U_row", {"SB200_GEMM_BT": "0"}),
         ("no_fused_solves", {"SB200_TRSM_FUSED": "0"})]
MIXED = [("default", {}), ("tile_fused", {"SB200_TILE_FUSED": "1"})]
GEMM = [("default", {}), ("untransposed_B_panel", {"SB200_GEMM_BT": "0"})]
GMIXED = [("default", {}), ("rows_in_shared_memory", {"SB200_PANEL_V4": "0"})]


def one(routine, n, nb):
    import torch
    import slate_b200.host as sl
    torch.cuda.set_device(0)
    reps = 3
    if routine == "potrf":
        fl = n ** 3 / 3 + n ** 2 / 2 + n / 6
        A0 = sl.HermitianMatrix(n, nb).generate("rand_dominant", 42); A = sl.HermitianMatrix(n, nb)
        fn = lambda: sl.potrf(A)
    elif routine == "getrf":
        fl = 2 * n ** 3 / 3 - n * n / 2 + 5 * n / 6
        A0 = sl.Matrix(n, n, nb).generate("rand", 42); A = sl.Matrix(n, n, nb)
        fn = lambda: sl.getrf(A)
    elif routine == "gemm":
        fl = 2.0 * n ** 3
        A0 = sl.Matrix(n, n, nb).generate("rand", 3); A = sl.Matrix(n, n, nb)
        Ga = sl.Matrix(n, n, nb).generate("rand", 1); Gb = sl.Matrix(n, n, nb).generate("rand", 2)
        fn = lambda: sl.gemm(3.1, Ga, Gb, 2.7, A)
    elif routine == "gesv_mixed":
        fl = 2 * n ** 3 / 3
        A0 = sl.Matrix(n, n, nb).generate("rand", 42); A = sl.Matrix(n, n, nb)
        B = sl.Matrix(n, 10, nb).generate("rand", 43); X = sl.Matrix(n, 10, nb)
        fn = lambda: sl.gesv_mixed(A, B, X)
    else:                                                   # posv_mixed: the FP32 factor is the chain there
        fl = n ** 3 / 3
        A0 = sl.HermitianMatrix(n, nb).generate("rand_dominant", 42); A = sl.HermitianMatrix(n, nb)
        B = sl.Matrix(n, 10, nb).generate("rand", 43); X = sl.Matrix(n, 10, nb)
        fn = lambda: sl.posv_mixed(A, B, X)
    best, tms = 1e30, None
    for r in range(reps):
        A.copy_from(A0)
        os.environ["SB200_PHASES"] = "1" if r == reps - 1 else "0"
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        out = fn()
        ev1.record(); torch.cuda.synchronize()
        mixed = routine in ("posv_mixed", "gesv_mixed")
        ms = A.last_driver_ms if not mixed else ev0.elapsed_time(ev1)
        best = min(best, ms)
        if mixed:
            tms = out[-1]
    print(json.dumps({"routine": routine, "n": n, "nb": nb, "best_ms": round(best, 2), "tflops": round(fl / best / 1e9, 2),
                      "panel_ms": round(A.last_panel_ms, 1) if not mixed else None, "timers": tms,
                      "env": {k: v for k, v in os.environ.items() if k.startswith("SB200_") and k != "SB200_PHASES"}}), flush=True)


if __name__ == "__main__":
    if sys.argv[1] == "one":
        one(sys.argv[2], int(sys.argv[3]), int(sys.argv[4]))
        sys.exit(0)
    routine = sys.argv[1]
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 32768
    nb = int(sys.argv[3]) if len(sys.argv) > 3 else 512
    only = [t for t in os.environ.get("SB200_VARIANTS", "").split(",") if t]
    for tag, env in {"potrf": POTRF, "getrf": GETRF, "gemm": GEMM, "posv_mixed": MIXED, "gesv_mixed": GMIXED}[routine]:
        if only and tag not in only:
            continue
        e = dict(os.environ); e.update(env)
        print(f"## {routine} {tag}", flush=True)
        sys.stderr.write(f"## {routine} {tag}\n"); sys.stderr.flush()
        try:
            subprocess.run([sys.executable, os.path.abspath(__file__), "one", routine, str(n), str(nb)], env=e, timeout=240, check=False)
        except subprocess.TimeoutExpired:
            print(f"{routine} {tag}: TIMEOUT (240 s)", flush=True)
