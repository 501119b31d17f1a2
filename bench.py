#!/usr/bin/env python
"""bench.py -- headline benchmark of the hot path (see BASELINE.json / DESIGN.md section Measurement).

    python bench.py --gpus N --steps K --warmup W            # our CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # reference CPU path (HostTask)

A "step" is ONE full factorisation (default routine: dpotrf, Cholesky) of a fresh seeded
matrix that is already resident in HBM when the timed region starts (`value`), and the same
through the public API from pinned HOST buffers, H2D + D2H inside the timed region (`e2e`).

N = 1 workload: BASELINE.json configs[1]  "dpotrf n=32768 nb=512 on 1 B200".
N > 1: the same routine at the n the headline metric is quoted on (n = 65536, fixed total work
for N = 2, 4, 8 -> "strong"), 2-D block-cyclic over a p x q grid of N ranks (one process per
GPU, NCCL panel broadcast).  The metric is a rate (TFLOP/s), so N = 1 at n = 32768 is comparable.

Prints exactly one JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)


def flops(routine: str, n: int) -> float:
    """The reference tester's own flop counts (lapackpp/include/lapack/flops.hh:27-60,
    blaspp/include/blas/flops.hh:100-104)."""
    n = float(n)
    if routine == "potrf":
        return (n ** 3 / 6 + n ** 2 / 2 + n / 3) + (n ** 3 / 6 - n / 6)
    if routine == "getrf":
        return (n ** 3 / 2 - n ** 3 / 6 + n * n / 2 - n * n / 2 + 2 * n / 3) + (n ** 3 / 2 - n ** 3 / 6 - n * n / 2 + n / 6)
    if routine == "gemm":
        return 2.0 * n ** 3
    raise ValueError(routine)


def _switches() -> dict:
    """Library switches in effect (SB200_* environment): empty for the default paths, so that a line measured with an
    opt-in candidate turned on says so."""
    return {k: v for k, v in sorted(os.environ.items())
            if k.startswith("SB200_") and k not in ("SB200_PHASES", "SB200_REFERENCE", "SB200_RUN_UNVALIDATED")}


def default_n(routine: str, ngpus: int) -> int:
    # N = 1: the single-GPU configuration BASELINE.json names (configs[1], n = 32768);
    # N > 1: the n the headline metric is quoted on (n = 65536), fixed total work for N = 2, 4, 8.
    if routine == "gemm":
        return 16384 if ngpus == 1 else 32768
    return 32768 if ngpus == 1 else 65536


class ClockSampler(threading.Thread):
    """Samples nvidia-smi SM clocks / throttle reasons during the timed region."""

    def __init__(self, index: int, period: float = 0.2):
        super().__init__(daemon=True)
        self.index = index
        self.period = period
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._halt = threading.Event()

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._halt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(",")]
                self.samples.append(float(f[0]))
                self.max_mhz = float(f[1])
                for nm, v in zip(names, f[2:]):
                    if v.lower().startswith("active"):
                        self.reasons.add(nm)
            except Exception:
                pass
            self._halt.wait(self.period)

    def stop(self):
        self._halt.set()
        self.join(timeout=10)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


def cpu_reference_run(routine: str, n: int, nb: int, threads: int):
    """Time the UNMODIFIED reference's HostTask path (oracle/_ref/ref_dump) on the host cores.
    Falls back to the numpy restatement (kind 'port') only if oracle/_ref is absent."""
    exe = os.path.join(HERE, "oracle", "_ref", "ref_dump")
    if os.path.exists(exe):
        env = dict(os.environ, OMP_NUM_THREADS=str(threads), OPENBLAS_NUM_THREADS="1")
        out = subprocess.run([exe, routine, "d", str(n), str(nb), "42", "43", "44", "/tmp/_sb200_ref", "dump=0"],
                             capture_output=True, text=True, env=env, timeout=1800)
        line = [l for l in out.stdout.splitlines() if l.startswith("{")]
        if not line:
            raise RuntimeError("ref_dump failed: " + out.stderr[-400:])
        r = json.loads(line[-1])
        return r["seconds"], "reference"
    import numpy as np
    from oracle import slate_oracle as o
    if routine == "potrf":
        G = o.generate("rand_dominant", n, n, 42)
        A = np.tril(G) + np.tril(G, -1).T
        t0 = time.time(); o.potrf(A, nb); return time.time() - t0, "port"
    if routine == "getrf":
        A = o.generate("rand", n, n, 42)
        t0 = time.time(); o.getrf(A, nb); return time.time() - t0, "port"
    A = o.generate("rand", n, n, 42); B = o.generate("rand", n, n, 43); C = o.generate("rand", n, n, 44)
    t0 = time.time(); o.gemm(3.1, A, B, 2.7, C, nb); return time.time() - t0, "port"


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path, all host threads,
    on a bounded sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    threads = os.cpu_count() or 1
    n = args.ref_n
    for _ in range(args.warmup if args.warmup < 2 else 1):
        cpu_reference_run(args.routine, n, args.nb, threads)
    secs, kind = [], "reference"
    for _ in range(args.steps):
        s, kind = cpu_reference_run(args.routine, n, args.nb, threads)
        secs.append(s)
    ms = 1e3 * sum(secs) / len(secs)
    if args.routine in ("posv_mixed", "gesv_mixed"):
        val = extra_flops(args.routine, n, args.nrhs) / (ms * 1e-3) / 1e12
    elif args.routine in EXTRA_ROUTINES:
        print(json.dumps({"impl": "reference", "unavailable": f"no CPU reference leg for --routine {args.routine}"}))
        return 0
    else:
        val = flops(args.routine, n) / (ms * 1e-3) / 1e12
    sample = f"d{args.routine} n={n} nb={args.nb} Target::HostTask (OpenMP tasks + OpenBLAS), {threads} threads"
    line = {
        "impl": "reference", "metric": f"d{args.routine} TFLOP/s", "value": val, "unit": "TFLOP/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic (reference matgen Philox rand/rand_dominant, seed 42)",
        "config": {"workload": sample, "routine": args.routine, "n": n, "nb": args.nb},
        "cpu_baseline": {"value": val, "unit": "TFLOP/s", "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": val, "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------------
# Secondary bench lines (not the driver's default): BASELINE configs[3] (zgemm / zherk), configs[4]
# (mixed-precision solves) and the HBM-bound tile kernels.
#   python bench.py --routine zgemm|zherk|zpotrf|posv_mixed|gesv_mixed|tileops [--n N] [--gpus N]
# ------------------------------------------------------------------------------------------------
EXTRA_ROUTINES = ["zgemm", "zherk", "zpotrf", "posv_mixed", "gesv_mixed", "tileops"]


def extra_flops(routine: str, n: int, nrhs: int) -> float:
    n = float(n)
    if routine == "zgemm":
        return 8.0 * n ** 3                      # blaspp flops.hh: complex fma = 6 mul-flops + 2 add-flops
    if routine == "zherk":
        return 4.0 * n * n * (n + 1)             # n x n result, k = n: herk = 4 * k n (n+1) / ... real-flop count
    if routine == "zpotrf":
        return 4.0 * flops("potrf", int(n))
    if routine == "posv_mixed":
        return flops("potrf", int(n)) + 2.0 * n * n * nrhs      # lapack::Gflop::posv (FP64-equivalent work)
    if routine == "gesv_mixed":
        return flops("getrf", int(n)) + 2.0 * n * n * nrhs      # lapack::Gflop::gesv
    raise ValueError(routine)


def measured_peaks():
    try:
        return json.load(open(os.path.join(HERE, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


def fp64_peak_probe(lib, st):
    """FP64 DMMA.8x8x4 pipe peak measured live (MEASURED_PEAKS.json has no FP64 figure), TFLOP/s."""
    import ctypes
    import torch
    from slate_b200._lib import c_dbl, c_int, c_ptr
    lib.sb200_fp64_peak_probe.argtypes = [c_int, c_int, c_int, c_ptr, ctypes.POINTER(c_dbl), c_ptr]
    scratch = torch.zeros(16, dtype=torch.float64, device="cuda")
    pf = c_dbl(0)
    lib.sb200_fp64_peak_probe(0, 2000, 4, scratch.data_ptr(), ctypes.byref(pf), st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); lib.sb200_fp64_peak_probe(0, 40000, 4, scratch.data_ptr(), ctypes.byref(pf), st); e1.record()
    torch.cuda.synchronize()
    return pf.value / (e0.elapsed_time(e1) * 1e-3) / 1e12


def run_tileops(args, sl, lib, torch):
    """HBM-bound tile kernels (include/slate/internal/device.hh:92-281) over a batch of nb x nb FP64 tiles:
    achieved GB/s = algorithmic bytes (geadd 3mns, gescale 2mns, geset mns, gecopy mn(s+d), transpose 2mns,
    norms mns) / CUDA-event time on the launch stream.  The batch (>= 2 GiB per operand) is far larger than L2."""
    import ctypes
    from slate_b200._lib import c_dbl, c_flt, c_i64, c_int, c_ptr
    nb = args.nb
    batch = args.n or 1024                       # --n = number of tiles here
    st = torch.cuda.current_stream().cuda_stream
    te = nb * nb
    A = torch.rand(batch * te, dtype=torch.float64, device="cuda")
    B = torch.rand(batch * te, dtype=torch.float64, device="cuda")
    S = torch.empty(batch * te, dtype=torch.float32, device="cuda")
    vals = torch.zeros(batch * nb * 2, dtype=torch.float64, device="cuda")
    pa = torch.tensor([A.data_ptr() + 8 * te * i for i in range(batch)], dtype=torch.int64, device="cuda")
    pb = torch.tensor([B.data_ptr() + 8 * te * i for i in range(batch)], dtype=torch.int64, device="cuda")
    ps = torch.tensor([S.data_ptr() + 4 * te * i for i in range(batch)], dtype=torch.int64, device="cuda")

    def fn(name, argtypes):
        f = getattr(lib, name); f.argtypes = argtypes; f.restype = c_int
        return f
    geadd = fn("sb200_geadd_batched_d", [c_i64, c_i64, c_dbl, c_ptr, c_i64, c_dbl, c_ptr, c_i64, c_i64, c_ptr])
    gescale = fn("sb200_gescale_batched_d", [c_i64, c_i64, c_dbl, c_dbl, c_ptr, c_i64, c_i64, c_ptr])
    geset = fn("sb200_geset_batched_d", [c_i64, c_i64, c_dbl, c_dbl, c_ptr, c_i64, c_i64, c_ptr])
    gecopy = fn("sb200_gecopy_batched_ds", [c_i64, c_i64, c_ptr, c_i64, c_ptr, c_i64, c_i64, c_ptr])
    tzset = fn("sb200_tzset_batched_d", [c_int, c_i64, c_i64, c_dbl, c_dbl, c_ptr, c_i64, c_i64, c_ptr])
    transp = fn("sb200_transpose_batched_d", [c_int, c_i64, c_i64, c_ptr, c_i64, c_ptr, c_i64, c_i64, c_ptr])
    transp_ip = fn("sb200_transpose_inplace_batched_d", [c_int, c_i64, c_ptr, c_i64, c_i64, c_ptr])
    genorm = fn("sb200_genorm_batched_d", [c_int, c_int, c_i64, c_i64, c_ptr, c_i64, c_ptr, c_i64, c_i64, c_ptr])
    P = lambda t: t.data_ptr()
    mns = float(nb) * nb * 8 * batch
    ops = {
        "geadd": (lambda: geadd(nb, nb, 2.0, P(pa), nb, 0.5, P(pb), nb, batch, st), 3 * mns),
        "gescale": (lambda: gescale(nb, nb, 3.0, 3.0, P(pb), nb, batch, st), 2 * mns),
        "geset": (lambda: geset(nb, nb, 0.0, 1.0, P(pb), nb, batch, st), mns),
        "tzset_lower": (lambda: tzset(ord("L"), nb, nb, 0.0, 1.0, P(pb), nb, batch, st), mns * (nb + 1) / (2.0 * nb)),
        "gecopy_d2s": (lambda: gecopy(nb, nb, P(pa), nb, P(ps), nb, batch, st), 1.5 * mns),
        "transpose": (lambda: transp(0, nb, nb, P(pa), nb, P(pb), nb, batch, st), 2 * mns),
        "transpose_inplace": (lambda: transp_ip(0, nb, P(pb), nb, batch, st), 2 * mns),
        "genorm_max": (lambda: genorm(ord("M"), ord("M"), nb, nb, P(pa), nb, P(vals), 1, batch, st), mns),
        "genorm_one": (lambda: genorm(ord("O"), ord("M"), nb, nb, P(pa), nb, P(vals), nb, batch, st), mns),
        "genorm_inf": (lambda: genorm(ord("I"), ord("M"), nb, nb, P(pa), nb, P(vals), nb, batch, st), mns),
        "genorm_fro": (lambda: genorm(ord("F"), ord("M"), nb, nb, P(pa), nb, P(vals), 2, batch, st), mns),
    }
    res, launches0 = {}, lib.sb200_launch_count()
    sampler = ClockSampler(0); sampler.start()
    t_all0 = time.perf_counter()
    for name, (call, nbytes) in ops.items():
        for _ in range(max(args.warmup, 3)):
            rc = call()
            if rc != 0:
                raise SystemExit(f"bench.py tileops: {name} returned {rc}")
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            call()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        res[name] = {"ms": ms, "GBps": nbytes / (ms * 1e-3) / 1e9, "algorithmic_bytes": nbytes}
    wall_ms = (time.perf_counter() - t_all0) * 1e3
    clocks = sampler.stop()
    pk = measured_peaks()
    peak = pk.get("hbm_gbs")
    src = "MEASURED_PEAKS.json hbm_gbs (torch copy read+write)" if peak else "fallback 6550 GB/s (B200_PROFILING.md)"
    peak = peak or 6550.0
    total_bytes = sum(v["algorithmic_bytes"] for v in res.values())
    total_ms = sum(v["ms"] for v in res.values())
    for v in res.values():
        v["frac"] = v["GBps"] / peak
    line = {
        "metric": "tile kernels GB/s (geadd/gescale/geset/tzset/gecopy/transpose/norms, FP64 nb x nb tiles)",
        "value": total_bytes / (total_ms * 1e-3) / 1e9, "unit": "GB/s", "n_gpus": 1, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": total_ms, "wall_ms": wall_ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic (uniform random tiles on device)",
        "config": {"workload": f"{batch} tiles of {nb}x{nb} FP64 per operand ({batch * te * 8 / 2**30:.1f} GiB), batched launches",
                   "routine": "tileops", "nb": nb, "batch": batch,
                   "l2": "each operand (>= 2 GiB) is far larger than the 126 MB L2; no explicit flush"},
        "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": "tile_foreach_kernel (geadd)", "achieved": res["geadd"]["GBps"], "peak": peak,
                     "unit": "GB/s", "frac": res["geadd"]["GBps"] / peak, "traffic": None, "peak_source": src},
        "kernels": res, "cpu_baseline": None, "e2e": None,
        "gpu_launches": int(lib.sb200_launch_count() - launches0),
    }
    print(json.dumps(line), flush=True)
    return 0


def run_extra(args):
    import ctypes
    import numpy as np
    import torch
    import torch.distributed as dist
    import slate_b200.host as sl
    from slate_b200._lib import lib, check, c_dbl, c_ptr

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; slate_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    routine, nb, nrhs = args.routine, args.nb, args.nrhs
    if routine == "tileops":
        return run_tileops(args, sl, lib, torch)
    mixed = routine in ("posv_mixed", "gesv_mixed")
    if mixed and world > 1 and os.environ.get("SB200_RUN_UNVALIDATED") != "1":
        raise SystemExit("bench.py: the mixed-precision solve path on a p x q grid (csrc/solve_dist.cu) has not been validated "
                         "on GPUs yet; set SB200_RUN_UNVALIDATED=1 to run it")
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    n = args.n or ({"zgemm": 16384, "zherk": 16384, "zpotrf": 24576, "posv_mixed": 32768, "gesv_mixed": 32768}[routine]
                   if world == 1 else {"zgemm": 40960, "zherk": 40960, "zpotrf": 40960, "posv_mixed": 65536, "gesv_mixed": 65536}[routine])
    grid = sl.Grid.from_torch_distributed() if world > 1 else sl.Grid()
    st = torch.cuda.current_stream().cuda_stream

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    al, be = 3.141592653589793 + 1.414213562373095j, 2.718281828459045 + 1.732050807568877j
    timers, phase_log = {}, []
    if routine == "zgemm":
        Am = sl.Matrix(n, n, nb, grid, "z").generate("rand", 42)
        Bm = sl.Matrix(n, n, nb, grid, "z").generate("rand", 43)
        out = sl.Matrix(n, n, nb, grid, "z").generate("rand", 44)
        run = lambda: sl.gemm(al, Am, Bm, be, out)
        kernel = "gemm_zdmma_kernel (complex128 split-complex on the FP64 DMMA pipe)"
    elif routine == "zherk":
        Am = sl.Matrix(n, n, nb, grid, "z").generate("rand", 42)
        out = sl.HermitianMatrix(n, nb, grid, dtype="z").generate("rand", 44)
        run = lambda: sl.herk(al.real, Am, be.real, out)
        kernel = "gemm_zdmma_kernel (complex128 herk: triangle-masked diagonal tiles)"
    elif routine == "zpotrf":
        A0 = sl.HermitianMatrix(n, nb, grid, dtype="z").generate("rand_dominant", 42)
        out = sl.HermitianMatrix(n, nb, grid, dtype="z")

        def run():
            out.copy_from(A0)
            if sl.potrf(out) != 0:
                raise SystemExit("zpotrf: info != 0")
        kernel = "gemm_zdmma_kernel (complex128 trailing update of zpotrf)"
    else:
        herm = routine == "posv_mixed"
        Am = (sl.HermitianMatrix(n, nb, grid) if herm else sl.Matrix(n, n, nb, grid)).generate(
            "rand_dominant" if herm else "rand", 42)
        Bm = sl.Matrix(n, nrhs, nb, grid).generate("rand", 43)
        out = sl.Matrix(n, nrhs, nb, grid)

        def run():
            r = sl.posv_mixed(Am, Bm, out) if herm else sl.gesv_mixed(Am, Bm, out)
            if r[0] != 0 or r[1] < 0:
                raise SystemExit(f"{routine}: info={r[0]} iter={r[1]} (refinement did not converge)")
            timers.update(r[-1]); timers["iter"] = r[1]
            phase_log.append(dict(r[-1]))
        kernel = "gemm_tf32x3_kernel (tcgen05 kind::tf32 x3, FP32-emulated trailing update of the low-precision factor)"

    stats = (c_dbl * 4)()
    lib.sb200_last_driver_stats.argtypes = [c_ptr, ctypes.POINTER(c_dbl)]
    for _ in range(max(args.warmup, 3) if not mixed else 2):
        run()
    sampler = ClockSampler(local_rank, period=1.0 if mixed else 0.2); sampler.start()
    launches0 = lib.sb200_launch_count()
    barrier(); w0 = time.perf_counter()
    step_ms, trail_ms, trail_flops, trail_launches = [], 0.0, 0.0, 0.0
    del phase_log[:]
    for _ in range(args.steps):
        run()
        check(lib.sb200_last_driver_stats(out._h, stats))
        step_ms.append(stats[0]); trail_ms += stats[1]; trail_flops += stats[2]; trail_launches += stats[3]
    barrier(); w1 = time.perf_counter()
    launches = lib.sb200_launch_count() - launches0
    clocks = sampler.stop()
    timed_phases = {k: sum(p[k] for p in phase_log) / len(phase_log) for k in phase_log[0]} if phase_log else {}
    t = torch.tensor([sum(step_ms) / len(step_ms), (w1 - w0) * 1e3 / args.steps], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = float(t[0])
    fl = extra_flops(routine, n, nrhs)
    value = fl / (ms_per_step * 1e-3) / 1e12
    achieved = trail_flops / (trail_ms * 1e-3) / 1e12 if trail_ms > 0 else 0.0
    if mixed:
        pk = measured_peaks()
        bf16 = pk.get("bf16_tflops_sustained") or pk.get("bf16_tflops") or 1500.0
        peak = bf16 / 2.0 / 3.0
        psrc = ("FP32-equivalent tensor peak = dense TF32 rate (half the measured bf16 rate in MEASURED_PEAKS.json: "
                f"{bf16:.0f} TF/s) / 3 tcgen05 MMAs per emulated FP32 product")
    else:
        peak = fp64_peak_probe(lib, st)
        psrc = "FP64 DMMA.8x8x4 probe measured live in this run (sb200_fp64_peak_probe)"
    roofline = {"bound": "tensor", "kernel": kernel, "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                "frac": achieved / peak if peak else None, "traffic": None, "peak_source": psrc,
                "launches_timed": int(trail_launches), "trailing_ms_per_step": trail_ms / args.steps}

    # e2e: operands from pinned HOST memory, result back to the host, inside the timed region
    e2e = None
    if not args.no_e2e:
        tdt = torch.complex128 if routine[0] == "z" else torch.float64
        esz = 16 if routine[0] == "z" else 8
        ins = [m for m in ((Am, Bm, out) if routine == "zgemm" else (Am, out) if routine == "zherk" else
                           (A0,) if routine == "zpotrf" else (Am, Bm))]
        dst = {id(A0): out} if routine == "zpotrf" else {}
        hosts = []
        for m in ins:
            h = torch.empty(m.local_tiles * nb * nb, dtype=tdt).pin_memory()
            m.to_host_local(h); hosts.append(h)
        res = torch.empty(out.local_tiles * nb * nb, dtype=tdt).pin_memory()
        e2e_ms = []
        for it in range(1 + min(args.steps, 2)):
            barrier(); t0 = time.perf_counter()
            for m, h in zip(ins, hosts):
                dst.get(id(m), m).from_host_local(h, sync=False)
            if routine == "zpotrf":
                if sl.potrf(out) != 0:
                    raise SystemExit("zpotrf: info != 0")
            else:
                run()
            out.to_host_local(res)
            barrier(); t1 = time.perf_counter()
            if it >= 1:
                e2e_ms.append((t1 - t0) * 1e3)
        em = sum(e2e_ms) / len(e2e_ms)
        h2d = sum(h.numel() for h in hosts) * esz
        d2h = res.numel() * esz
        te = torch.tensor([em, float(h2d), float(d2h)], dtype=torch.float64, device="cuda")
        if world > 1:
            tmax = te.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            tsum = te.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
            em, h2d, d2h = float(tmax[0]), float(tsum[1]), float(tsum[2])
        e2e = {"value": fl / (em * 1e-3) / 1e12, "unit": "TFLOP/s", "ms_per_step": em,
               "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "note": "pinned host tiles -> from_host_local -> driver -> to_host_local; bytes summed over ranks"}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and mixed:
        try:
            threads = os.cpu_count() or 1
            rn = min(args.ref_n, 8192)
            secs, kind = cpu_reference_run(routine, rn, nb, threads)
            cpu = {"value": extra_flops(routine, rn, nrhs) / secs / 1e12, "unit": "TFLOP/s", "cores": threads, "kind": kind,
                   "sample": f"d{routine} n={rn} nb={nb} nrhs={nrhs} Target::HostTask, one run, {secs:.2f} s"}
        except Exception as ex:   # noqa: BLE001
            cpu = {"value": None, "unit": "TFLOP/s", "cores": os.cpu_count(), "kind": "reference", "sample": f"failed: {ex}"}
    if rank == 0:
        name = routine if routine[0] == "z" else "d" + routine
        line = {
            "metric": f"{name} TFLOP/s" + (" (FP64-equivalent: lapack::Gflop of the FP64 solve / time of the mixed solve)" if mixed else ""),
            "value": value, "unit": "TFLOP/s", "n_gpus": world, "steps": args.steps,
            "warmup": 1 if mixed else max(args.warmup, 3), "ms_per_step": ms_per_step,
            "wall_ms_per_step": float(t[1]), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "c128" if routine[0] == "z" else "f32 factor (3xTF32 tensor cores) + f64 refinement",
            "data": "synthetic (reference matgen: Philox-2x64 rand_dominant/rand, seeds 42/43/44, generated on device)",
            "config": {"workload": f"{name} n={n} nb={nb}" + (f" nrhs={nrhs}" if mixed else "")
                                   + f", {grid.p}x{grid.q} block-cyclic grid over {world} B200",
                       "routine": routine, "n": n, "nb": nb, "grid": [grid.p, grid.q],
                       "l2": "operands (GiBs) are far larger than the 126 MB L2; no explicit flush",
                       "switches": _switches()},
            "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches),
        }
        line["step_ms"] = step_ms
        if mixed:
            timed_phases["iter"] = timers.get("iter")
            line["phases_ms"] = timed_phases        # mean over the timed steps (reference timer names)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--routine", default="potrf", choices=["potrf", "getrf", "gemm"] + EXTRA_ROUTINES)
    ap.add_argument("--nrhs", type=int, default=10)
    ap.add_argument("--n", "--size", dest="n", type=int, default=0,
                    help="matrix size (use --size under torchrun, whose own parser claims the prefix --n)")
    ap.add_argument("--nb", type=int, default=512)
    ap.add_argument("--ref-n", type=int, default=8192, help="bounded sample size for the CPU reference legs")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.routine in EXTRA_ROUTINES:
        return run_extra(args)

    import ctypes
    import numpy as np
    import torch
    import torch.distributed as dist
    import slate_b200.host as sl
    from slate_b200._lib import lib, check, c_dbl, c_int, c_ptr

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; slate_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torchrun --nproc-per-node {args.gpus}")
    routine, nb = args.routine, args.nb
    n = args.n or default_n(routine, world)
    grid = sl.Grid.from_torch_distributed() if world > 1 else sl.Grid()
    st = torch.cuda.current_stream().cuda_stream

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- operands (resident in HBM before the timed region); a pristine copy restores the input
    if routine == "potrf":
        A0 = sl.HermitianMatrix(n, nb, grid).generate("rand_dominant", 42)
        A = sl.HermitianMatrix(n, nb, grid)
        run = lambda: sl.potrf(A)
        restore = lambda: A.copy_from(A0)
        out = A
    elif routine == "getrf":
        A0 = sl.Matrix(n, n, nb, grid).generate("rand", 42)
        A = sl.Matrix(n, n, nb, grid)
        run = lambda: sl.getrf(A)[1]
        restore = lambda: A.copy_from(A0)
        out = A
    else:
        Am = sl.Matrix(n, n, nb, grid).generate("rand", 42)
        Bm = sl.Matrix(n, n, nb, grid).generate("rand", 43)
        A0 = sl.Matrix(n, n, nb, grid).generate("rand", 44)
        A = sl.Matrix(n, n, nb, grid)
        run = lambda: (sl.gemm(3.141592653589793, Am, Bm, 2.718281828459045, A), 0)[1]
        restore = lambda: A.copy_from(A0)
        out = A

    stats = (c_dbl * 4)()
    lib.sb200_last_driver_stats.argtypes = [c_ptr, ctypes.POINTER(c_dbl)]

    for _ in range(max(args.warmup, 3)):
        restore(); info = run()
        if info != 0:
            raise SystemExit(f"bench.py: {routine} returned info={info}")

    # ---- timed region: K steps, device time by CUDA events inside the driver (on its own
    #      streams), bracketed by barrier + synchronize; max over ranks
    sampler = ClockSampler(local_rank); sampler.start()
    launches0 = lib.sb200_launch_count()
    barrier(); w0 = time.perf_counter()
    step_ms, trail_ms, trail_flops, trail_launches, panel_ms = [], 0.0, 0.0, 0.0, 0.0
    for _ in range(args.steps):
        restore(); run()
        check(lib.sb200_last_driver_stats(out._h, stats))
        step_ms.append(stats[0]); trail_ms += stats[1]; trail_flops += stats[2]; trail_launches += stats[3]
        panel_ms += out.last_panel_ms
    barrier(); w1 = time.perf_counter()
    launches = lib.sb200_launch_count() - launches0
    clocks = sampler.stop()
    t = torch.tensor([sum(step_ms) / len(step_ms), (w1 - w0) * 1e3 / args.steps], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step, wall_ms_per_step = float(t[0]), float(t[1])
    fl = flops(routine, n)
    value = fl / (ms_per_step * 1e-3) / 1e12

    # ---- FP64 tensor peak measured in this run (MEASURED_PEAKS.json has no FP64 figure)
    lib.sb200_fp64_peak_probe.argtypes = [c_int, c_int, c_int, c_ptr, ctypes.POINTER(c_dbl), c_ptr]
    scratch = torch.zeros(16, dtype=torch.float64, device="cuda")
    pf = c_dbl(0)
    lib.sb200_fp64_peak_probe(0, 2000, 4, scratch.data_ptr(), ctypes.byref(pf), st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); lib.sb200_fp64_peak_probe(0, 40000, 4, scratch.data_ptr(), ctypes.byref(pf), st); e1.record()
    torch.cuda.synchronize()
    peak = pf.value / (e0.elapsed_time(e1) * 1e-3) / 1e12
    achieved = trail_flops / (trail_ms * 1e-3) / 1e12 if trail_ms > 0 else 0.0
    traffic = None
    tfile = os.path.join(HERE, "profiles", "gemm_dram_bytes_per_launch.json")
    if os.path.exists(tfile):
        try:
            traffic = json.load(open(tfile)).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    roofline = {"bound": "tensor", "kernel": "gemm_dmma_kernel (trailing-update batched tile GEMM/HERK, FP64 DMMA)",
                "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak if peak else None,
                "traffic": traffic,
                "peak_source": "FP64 DMMA.8x8x4 probe measured live in this run (sb200_fp64_peak_probe); "
                               "MEASURED_PEAKS.json carries no FP64 figure",
                "launches_timed": int(trail_launches),
                "trailing_ms_per_step": trail_ms / args.steps, "panel_stream_ms_per_step": panel_ms / args.steps,
                "whole_step_frac_of_peak": value / (world * peak) if peak else None}

    # ---- e2e: public API with HOST buffers (pinned), H2D + driver + D2H inside the timed region.
    #      Every rank moves ITS tiles (packed local tile storage, as Matrix::insertLocalTiles hands
    #      SLATE caller-owned memory); wall clock between barriers, max over ranks.
    e2e = None
    if not args.no_e2e:
        nelem = out.local_tiles * nb * nb
        host = torch.empty(nelem, dtype=torch.float64).pin_memory()
        res = torch.empty(nelem, dtype=torch.float64).pin_memory()
        A0.to_host_local(host)            # setup (untimed): host copy of the seeded input
        e2e_ms = []
        # opt-in (round-2 candidate, not yet run): D2H of finished block columns overlapped inside the driver
        overlap_d2h = routine == "potrf" and os.environ.get("SB200_E2E_OVERLAP") == "1"
        # opt-in (round-2 candidate, not yet run; one rank): the input streams in by chunks of block columns as well
        overlap_both = routine == "potrf" and world == 1 and os.environ.get("SB200_E2E_OVERLAP") == "2"
        for it in range(1 + args.steps):
            barrier(); t0 = time.perf_counter()
            if overlap_both:
                sl.potrf(A, in_local=host, out_local=res)
                barrier(); t1 = time.perf_counter()
                if it >= 1:
                    e2e_ms.append((t1 - t0) * 1e3)
                continue
            A.from_host_local(host, sync=False)
            if overlap_d2h:
                sl.potrf(A, out_local=res)          # finished block columns stream to the host while it factors
            else:
                run()
                A.to_host_local(res)
            barrier(); t1 = time.perf_counter()
            if it >= 1:
                e2e_ms.append((t1 - t0) * 1e3)
        te = torch.tensor([sum(e2e_ms) / len(e2e_ms), float(nelem * 8)], dtype=torch.float64, device="cuda")
        if world > 1:
            tmax = te.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            tsum = te.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
            em, nbytes = float(tmax[0]), float(tsum[1])
        else:
            em, nbytes = float(te[0]), float(te[1])
        e2e = {"value": fl / (em * 1e-3) / 1e12, "unit": "TFLOP/s", "ms_per_step": em,
               "h2d_bytes_per_step": int(nbytes), "d2h_bytes_per_step": int(nbytes),
               "note": "pinned host tiles -> Matrix.from_host_local -> driver -> Matrix.to_host_local on every rank; "
                       "bytes summed over ranks, time = max over ranks"}
        del host, res

    # ---- CPU baseline: reference HostTask on this box's cores, bounded sample (rank 0, N = 1)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            threads = os.cpu_count() or 1
            secs, kind = cpu_reference_run(routine, args.ref_n, nb, threads)
            cpu = {"value": flops(routine, args.ref_n) / secs / 1e12, "unit": "TFLOP/s", "cores": threads, "kind": kind,
                   "sample": f"d{routine} n={args.ref_n} nb={nb} Target::HostTask, one run, {secs:.2f} s"}
        except Exception as ex:   # noqa: BLE001
            cpu = {"value": None, "unit": "TFLOP/s", "cores": os.cpu_count(), "kind": "reference", "sample": f"failed: {ex}"}

    if rank == 0:
        p, q = grid.p, grid.q
        line = {
            "metric": f"d{routine} TFLOP/s", "value": value, "unit": "TFLOP/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step,
            "wall_ms_per_step": wall_ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64",
            "data": "synthetic (reference matgen: Philox-2x64 rand_dominant/rand, seed 42, generated on device)",
            "config": {"workload": f"d{routine} n={n} nb={nb} lookahead=1, {p}x{q} block-cyclic grid over {world} B200",
                       "routine": routine, "n": n, "nb": nb, "grid": [p, q],
                       "l2": "inputs (>= 4 GiB per step) are far larger than the 126 MB L2; no explicit flush",
                       "switches": _switches()},
            "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e,
            "gpu_launches": int(launches),
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
