#!/usr/bin/env python
"""bench.py -- headline benchmark of the hot path (see BASELINE.json / DESIGN.md section Measurement).

    python bench.py --gpus N --steps K --warmup W            # our CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # reference CPU path (HostTask)

A "step" is ONE full factorisation (default routine: dpotrf, Cholesky) of a fresh seeded
matrix that is already resident in HBM when the timed region starts (`value`), and the same
through the public API from pinned HOST buffers, H2D + D2H inside the timed region (`e2e`).

Workload at every N: the n the headline metric is quoted on (BASELINE.json: dpotrf/dgetrf/dgemm at
n = 65536 on 1/2/4/8 B200; it fits one GPU), nb = 512, 2-D block-cyclic over a p x q grid of N ranks
(one process per GPU, NCCL panel broadcasts) -> fixed total work, "scaling": "strong".  The default
run times dpotrf as the line's `value` and dgetrf / dgemm at the same n as `also` sub-records, each
with an untimed probe-vector residual `check` (the tester's tolerances).  BASELINE configs[1]
(n = 32768) is `--size 32768`.  `value` is wall clock between barriers (driver entry to return, the
restore copy of the input included and reported); `device_ms_per_step` is the in-driver event time.

Prints exactly one JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import math
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
            if k.startswith("SB200_") and k not in ("SB200_PHASES", "SB200_REFERENCE", "SB200_HOST_TIMES")}


def default_n(routine: str, ngpus: int) -> int:
    # the n the headline metric is quoted on (BASELINE.json: "dpotrf/dgetrf/dgemm TFLOP/s at n=65536, 1/2/4/8 B200"),
    # the SAME at every N, so "scaling": "strong" is literally true.  n = 65536 fits one B200 (potrf: 2 x 17 GB, getrf:
    # 2 x 34 GB, gemm: 3 x 34 GB of the 180 GB); BASELINE configs[1] (n = 32768) is `--size 32768`.
    return 65536


class ClockSampler(threading.Thread):
    """Samples SM clocks / throttle reasons during the timed region.  In-process NVML (nvidia_ml_py) queries: spawning
    `nvidia-smi` five times a second was measured (r2b) to stall the CUDA API calls of the launch loop by 40-280 ms per
    step.  Falls back to `nvidia-smi` once a second when NVML cannot be loaded."""

    def __init__(self, index: int, period: float = 0.25):
        super().__init__(daemon=True)
        self.index = index
        self.period = period
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._halt = threading.Event()
        self._nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = index
            if vis:
                ids = [x for x in vis.split(",") if x.strip() != ""]
                if index < len(ids) and ids[index].strip().isdigit():
                    phys = int(ids[index])
            self._h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
            self._nvml = pynvml
        except Exception:
            self._nvml = None
            self.period = max(period, 1.0)

    def _sample_nvml(self):
        n = self._nvml
        self.samples.append(float(n.nvmlDeviceGetClockInfo(self._h, n.NVML_CLOCK_SM)))
        try:
            r = n.nvmlDeviceGetCurrentClocksEventReasons(self._h)
        except Exception:
            r = n.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
        for nm, bit in (("hw_slowdown", 0x8), ("sw_power_cap", 0x4), ("sw_thermal_slowdown", 0x20),
                        ("hw_thermal_slowdown", 0x40)):
            if r & bit:
                self.reasons.add(nm)

    def _sample_smi(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                              "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout
        f = [x.strip() for x in out.strip().split(",")]
        self.samples.append(float(f[0]))
        self.max_mhz = float(f[1])
        for nm, v in zip(names, f[2:]):
            if v.lower().startswith("active"):
                self.reasons.add(nm)

    def run(self):
        while not self._halt.is_set():
            try:
                if self._nvml:
                    self._sample_nvml()
                else:
                    self._sample_smi()
            except Exception:
                pass
            self._halt.wait(self.period)

    def stop(self):
        self._halt.set()
        self.join(timeout=10)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s), "source": "nvml" if self._nvml else "nvidia-smi"}


# bench routine -> (ref_dump routine, type, extra keys) of the reference's own HostTask run
REF_ROUTINE = {"zgemm": ("gemm", "z", []), "zherk": ("herk", "z", []), "zpotrf": ("potrf", "z", []), "zgetrf": ("getrf", "z", []),
               "getrf_tntpiv": ("getrf", "d", ["method=calu"])}


def cpu_reference_run(routine: str, n: int, nb: int, threads: int):
    """Time the UNMODIFIED reference's HostTask path (oracle/_ref/ref_dump) on the host cores.
    Falls back to the numpy restatement (kind 'port') only if oracle/_ref is absent."""
    exe = os.path.join(HERE, "oracle", "_ref", "ref_dump")
    if os.path.exists(exe):
        env = dict(os.environ, OMP_NUM_THREADS=str(threads), OPENBLAS_NUM_THREADS="1")
        rr, rt, rkv = REF_ROUTINE.get(routine, (routine, "d", []))
        out = subprocess.run([exe, rr, rt, str(n), str(nb), "42", "43", "44", "/tmp/_sb200_ref", "dump=0"] + rkv,
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


REF_RATE = {"potrf": 0.65e12, "getrf": 0.45e12, "gemm": 0.8e12}     # HostTask on 16 host threads, as measured in round 1 / 2
REF_PROBE_N = 8192                                                  # size of the rate probe of --impl reference


def pick_ref_n(routine, runs, budget_s=150.0, rate=None):
    """Largest reference sample size whose `runs` executions fit the time budget: the HostTask rate still rises with n,
    so the closer the sample is to the metric's n = 65536 the less the reference arm is under-estimated."""
    r = routine if routine in ("potrf", "getrf", "gemm") else "getrf"
    rate = rate or REF_RATE[r]
    for n in (32768, 24576, 16384, 8192):
        if runs * flops(r, n) / rate <= budget_s:
            return n
    return 8192


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path, all host threads,
    on a bounded sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    threads = os.cpu_count() or 1
    if args.routine == "tileops":
        print(json.dumps({"impl": "reference", "unavailable": "no CPU reference leg for --routine tileops"}))
        return 0
    probe = None
    if not args.ref_n and args.routine in REF_RATE:
        # the tabulated rates are from one box; measure this box's (one short run: the HostTask rate only rises with n,
        # so a size chosen with the probe's rate finishes inside the bound)
        try:
            s0, _ = cpu_reference_run(args.routine, REF_PROBE_N, args.nb, threads)
            probe = flops(args.routine, REF_PROBE_N) / s0
        except Exception:   # noqa: BLE001
            probe = None
    n = args.ref_n or pick_ref_n(args.routine, args.steps + 1, rate=probe)
    if not args.ref_n and args.routine[0] == "z":
        n = min(n, 8192)                      # four times the real flops per element: keep the sample within the time bound
    for _ in range(args.warmup if args.warmup < 2 else 1):
        cpu_reference_run(args.routine, n, args.nb, threads)
    secs, kind = [], "reference"
    for _ in range(args.steps):
        s, kind = cpu_reference_run(args.routine, n, args.nb, threads)
        secs.append(s)
    ms = 1e3 * sum(secs) / len(secs)
    if args.routine in EXTRA_ROUTINES:
        val = extra_flops(args.routine, n, args.nrhs) / (ms * 1e-3) / 1e12
    else:
        val = flops(args.routine, n) / (ms * 1e-3) / 1e12
    n_full = args.n or default_n(args.routine, args.gpus)
    p = int(math.floor(math.sqrt(args.gpus)))
    while args.gpus % p:
        p -= 1
    q = args.gpus // p
    sample = (f"d{args.routine} n={n} nb={args.nb} Target::HostTask (OpenMP tasks + OpenBLAS), {threads} host threads: a bounded "
              f"sample of the n={n_full} workload (same generator, seed and tile size; the full size takes minutes per step "
              f"on the host) -- the metric is a rate")
    line = {
        "impl": "reference", "metric": f"{args.routine if args.routine[0] == 'z' else 'd' + args.routine} TFLOP/s", "value": val, "unit": "TFLOP/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "c128" if args.routine[0] == "z" else "f64",
        "data": "synthetic (reference matgen Philox rand/rand_dominant, seed 42)",
        "config": workload_config(args.routine, n_full, args.nb, p, q, args.gpus),
        "sample_n": n, "sample_rate_probe_tflops": (probe / 1e12 if probe else None),
        "cpu_baseline": {"value": val, "unit": "TFLOP/s", "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": val, "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------------
# Secondary bench lines (not the driver's default): BASELINE configs[3] (zgemm / zherk), configs[4]
# (mixed-precision solves) and the HBM-bound tile kernels.
#   python bench.py --routine zgemm|zherk|zpotrf|zgetrf|getrf_tntpiv|posv_mixed|gesv_mixed|tileops [--n N] [--gpus N]
# ------------------------------------------------------------------------------------------------
EXTRA_ROUTINES = ["zgemm", "zherk", "zpotrf", "zgetrf", "getrf_tntpiv", "posv_mixed", "gesv_mixed", "tileops"]


def extra_flops(routine: str, n: int, nrhs: int) -> float:
    n = float(n)
    if routine == "zgemm":
        return 8.0 * n ** 3                      # blaspp flops.hh: complex fma = 6 mul-flops + 2 add-flops
    if routine == "zherk":
        return 4.0 * n * n * (n + 1)             # n x n result, k = n: herk = 4 * k n (n+1) / ... real-flop count
    if routine == "zpotrf":
        return 4.0 * flops("potrf", int(n))
    if routine == "zgetrf":
        return 4.0 * flops("getrf", int(n))      # complex LU (1 x 1 grid)
    if routine == "getrf_tntpiv":
        return flops("getrf", int(n))            # CALU: lapack::Gflop::getrf, as the tester reports it
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
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    n = args.n or ({"zgemm": 16384, "zherk": 16384, "zpotrf": 24576, "zgetrf": 16384, "getrf_tntpiv": 32768,
                    "posv_mixed": 32768, "gesv_mixed": 32768}[routine]
                   if world == 1 else {"zgemm": 40960, "zherk": 40960, "zpotrf": 40960, "zgetrf": 40960, "getrf_tntpiv": 65536,
                                       "posv_mixed": 65536, "gesv_mixed": 65536}[routine])
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
    elif routine in ("zgetrf", "getrf_tntpiv"):
        dt = "z" if routine == "zgetrf" else "d"
        A0 = sl.Matrix(n, n, nb, grid, dt).generate("rand", 42)
        out = sl.Matrix(n, n, nb, grid, dt)

        def run():
            out.copy_from(A0)
            _, info = sl.getrf(out) if routine == "zgetrf" else sl.getrf_tntpiv(out)
            if info != 0:
                raise SystemExit(f"{routine}: info != 0")
        kernel = ("gemm_zdmma_kernel (complex128 trailing update of zgetrf; cooperative complex panel)" if routine == "zgetrf"
                  else "gemm_dmma_kernel (FP64 trailing update of the LU with tournament pivoting)")
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

    # untimed check of a fresh factorisation: the tester's LU check with one probe vector (test/test_gesv.cc:371-377)
    extra_check = None
    if routine in ("zgetrf", "getrf_tntpiv"):
        try:
            out.copy_from(A0)
            piv, _ = sl.getrf(out) if routine == "zgetrf" else sl.getrf_tntpiv(out)
            extra_check = sl.getrf_residual(A0, out, piv)
        except Exception as ex:   # noqa: BLE001  (the timed numbers above stand on their own)
            extra_check = {"error": str(ex)}

    # e2e: operands from pinned HOST memory, result back to the host, inside the timed region
    e2e = None
    if not args.no_e2e:
        tdt = torch.complex128 if routine[0] == "z" else torch.float64
        esz = 16 if routine[0] == "z" else 8
        fact = routine in ("zpotrf", "zgetrf", "getrf_tntpiv")          # factorisations: host tiles -> out, factored in place
        ins = [m for m in ((Am, Bm, out) if routine == "zgemm" else (Am, out) if routine == "zherk" else
                           (A0,) if fact else (Am, Bm))]
        dst = {id(A0): out} if fact else {}
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
            elif fact:
                if (sl.getrf(out) if routine == "zgetrf" else sl.getrf_tntpiv(out))[1] != 0:
                    raise SystemExit(f"{routine}: info != 0")
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
            rn = min(args.ref_n or 8192, 8192)
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
            "dtype": "c128" if routine[0] == "z" else "f64" if routine == "getrf_tntpiv" else "f32 factor (3xTF32 tensor cores) + f64 refinement",
            "data": "synthetic (reference matgen: Philox-2x64 rand_dominant/rand, seeds 42/43/44, generated on device)",
            "config": {"workload": f"{name} n={n} nb={nb}" + (f" nrhs={nrhs}" if mixed else "")
                                   + f", {grid.p}x{grid.q} block-cyclic grid over {world} B200",
                       "routine": routine, "n": n, "nb": nb, "grid": [grid.p, grid.q],
                       "l2": "operands (GiBs) are far larger than the 126 MB L2; no explicit flush",
                       "switches": _switches()},
            "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches),
        }
        if extra_check is not None:
            line["check"] = extra_check
        line["step_ms"] = step_ms
        if mixed:
            timed_phases["iter"] = timers.get("iter")
            line["phases_ms"] = timed_phases        # mean over the timed steps (reference timer names)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def _barrier(torch, dist, world):
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def _max_over_ranks(torch, dist, world, vals):
    t = torch.tensor(vals, dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(v) for v in t]


GEMM_VARIANT = {"potrf": "NT", "getrf": "NT", "gemm": "NT"}     # the gemm_dmma_kernel variant each routine's update runs


def measured_traffic(routine):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the trailing-update kernel, from the committed
    `ncu --set full` capture of the variant THIS routine runs (profiles/gemm_dram_bytes_per_launch.json); None if that
    variant was never captured."""
    try:
        d = json.load(open(os.path.join(HERE, "profiles", "gemm_dram_bytes_per_launch.json")))
        v = d.get("variants", {}).get(f"{routine}:{GEMM_VARIANT[routine]}")
        return (v or {}).get("dram_bytes_per_launch"), (v or {}).get("algorithmic_bytes_per_launch"), (v or {}).get("source")
    except Exception:
        return None, None, None


def run_factor_bench(routine, n, nb, grid, world, steps, warmup, env, with_clocks=False, local_rank=0):
    """One routine at one size: warm-up, K timed steps bracketed by barrier + synchronize (wall clock, max over ranks;
    the restore copy of the input is inside the bracket and reported), device-event times from the driver, the trailing
    kernel's roofline numbers, and an untimed probe-vector residual check of the last result."""
    import ctypes
    torch, dist, sl, lib, check_, c_dbl, c_ptr = (env[k] for k in ("torch", "dist", "sl", "lib", "check", "c_dbl", "c_ptr"))
    al, be = 3.141592653589793, 2.718281828459045
    pivots = [None]
    if routine == "potrf":
        A0 = sl.HermitianMatrix(n, nb, grid).generate("rand_dominant", 42)
        A = sl.HermitianMatrix(n, nb, grid)
        run = lambda: sl.potrf(A)
        restore = lambda: A.copy_from(A0)
        mats = [A0, A]
    elif routine == "getrf":
        A0 = sl.Matrix(n, n, nb, grid).generate("rand", 42)
        A = sl.Matrix(n, n, nb, grid)

        def run():
            pivots[0], info = sl.getrf(A)
            return info
        restore = lambda: A.copy_from(A0)
        mats = [A0, A]
    else:
        Am = sl.Matrix(n, n, nb, grid).generate("rand", 42)
        Bm = sl.Matrix(n, n, nb, grid).generate("rand", 43)
        A = sl.Matrix(n, n, nb, grid).generate("rand", 44)
        A0 = None
        run = lambda: (sl.gemm(al, Am, Bm, be, A), 0)[1]
        restore = lambda: None          # C keeps accumulating (alpha A B + beta C); the check below runs on a fresh C
        mats = [Am, Bm, A]
    out = A
    stats = (c_dbl * 4)()
    lib.sb200_last_driver_stats.argtypes = [c_ptr, ctypes.POINTER(c_dbl)]
    for _ in range(warmup):
        restore(); info = run()
        if info != 0:
            raise SystemExit(f"bench.py: {routine} returned info={info}")
    sampler = None
    if with_clocks:
        sampler = ClockSampler(local_rank); sampler.start()
    launches0 = lib.sb200_launch_count()
    dev_ms, trail_ms, trail_flops, trail_launches, panel_ms = [], 0.0, 0.0, 0.0, 0.0
    _barrier(torch, dist, world); w0 = time.perf_counter()
    for _ in range(steps):
        restore(); run()
        check_(lib.sb200_last_driver_stats(out._h, stats))
        dev_ms.append(stats[0]); trail_ms += stats[1]; trail_flops += stats[2]; trail_launches += stats[3]
        panel_ms += out.last_panel_ms
    _barrier(torch, dist, world); w1 = time.perf_counter()
    launches = lib.sb200_launch_count() - launches0
    clocks = sampler.stop() if sampler else None
    # restore cost (inside the bracket above), measured once
    _barrier(torch, dist, world); r0 = time.perf_counter(); restore(); _barrier(torch, dist, world)
    restore_ms = (time.perf_counter() - r0) * 1e3
    wall_ms, device_ms, restore_ms = _max_over_ranks(torch, dist, world, [(w1 - w0) * 1e3 / steps, sum(dev_ms) / len(dev_ms), restore_ms])
    fl = flops(routine, n)
    rec = {"value": fl / (wall_ms * 1e-3) / 1e12, "ms_per_step": wall_ms, "device_ms_per_step": device_ms,
           "restore_ms_per_step": restore_ms if routine != "gemm" else 0.0, "steps": steps, "warmup": warmup, "n": n, "nb": nb,
           "trail_ms_per_step": trail_ms / steps, "trail_flops_per_step": trail_flops / steps,
           "trail_launches": int(trail_launches), "panel_stream_ms_per_step": panel_ms / steps,
           "gpu_launches": int(launches), "clocks": clocks}
    # ---- untimed check of a fresh result (probe-vector residual; tolerance = the tester's)
    if routine == "potrf":
        restore(); run()
        rec["check"] = sl.potrf_residual(A0, A)
    elif routine == "getrf":
        restore(); run()
        rec["check"] = sl.getrf_residual(A0, A, pivots[0])
    else:
        A.generate("rand", 44)
        x = sl._probe_vector(n, A.dtype, 5)
        c0x = sl.probe_mv(A, x, "G")
        run()
        rec["check"] = sl.gemm_residual(al, Am, Bm, be, c0x, A, x)
    return rec, mats, (A0, A)


def e2e_potrf(n, nb, grid, world, steps, env, mats):
    """The same metric through the public API with HOST buffers: pinned host tiles -> device -> potrf -> host, all inside
    the timed region (wall clock between barriers, max over ranks).  Mode 2 (one rank): the input streams in by chunks
    of block columns and finished block columns stream out while the factorisation runs (sl.potrf(in_local=, out_local=));
    mode 1: the output streams out; mode 0: copy, factor, copy."""
    torch, dist, sl = env["torch"], env["dist"], env["sl"]
    A0, A = mats
    mode = os.environ.get("SB200_E2E_OVERLAP")
    mode = int(mode) if mode is not None else (2 if world == 1 else 1)
    nelem = A.local_tiles * nb * nb
    host = torch.empty(nelem, dtype=torch.float64).pin_memory()
    res = torch.empty(nelem, dtype=torch.float64).pin_memory()
    A0.to_host_local(host)            # setup (untimed): host copy of the seeded input
    e2e_ms = []
    for it in range(1 + steps):
        _barrier(torch, dist, world); t0 = time.perf_counter()
        if mode == 2 and world == 1:
            info = sl.potrf(A, in_local=host, out_local=res)
        else:
            A.from_host_local(host, sync=False)
            if mode == 1:
                info = sl.potrf(A, out_local=res)
            else:
                info = sl.potrf(A)
                A.to_host_local(res)
        _barrier(torch, dist, world); t1 = time.perf_counter()
        if info != 0:
            raise SystemExit(f"bench.py: e2e potrf returned info={info}")
        if it >= 1:
            e2e_ms.append((t1 - t0) * 1e3)
    # the streamed result is the factor the resident path produces (bitwise on one rank)
    ok = None
    if world == 1:
        A.copy_from(A0); sl.potrf(A)
        chk = torch.empty(nelem, dtype=torch.float64).pin_memory()
        A.to_host_local(chk)
        ok = bool(torch.equal(chk, res))
        del chk
    te = torch.tensor([sum(e2e_ms) / len(e2e_ms)], dtype=torch.float64, device="cuda")
    tb = torch.tensor([float(nelem * 8)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
        dist.all_reduce(tb, op=dist.ReduceOp.SUM)
    em, nbytes = float(te[0]), float(tb[0])
    del host, res
    return {"value": flops("potrf", n) / (em * 1e-3) / 1e12, "unit": "TFLOP/s", "ms_per_step": em,
            "h2d_bytes_per_step": int(nbytes), "d2h_bytes_per_step": int(nbytes), "overlap_mode": mode,
            "matches_resident_result": ok,
            "note": "pinned host tiles -> device -> potrf -> pinned host tiles on every rank, copies inside the timed region "
                    "(mode 2: input and output stream by block columns while the factorisation runs; 1: output streams; 0: serial); "
                    "bytes summed over ranks, time = max over ranks"}


def workload_config(routine, n, nb, p, q, world):
    return {"workload": f"d{routine} n={n} nb={nb} lookahead=1, {p}x{q} block-cyclic grid over {world} B200",
            "routine": routine, "n": n, "nb": nb, "grid": [p, q],
            "l2": "operands (>= 16 GiB per step) are far larger than the 126 MB L2; no explicit flush"}


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
    ap.add_argument("--ref-n", type=int, default=0,
                    help="bounded sample size for the CPU reference legs (0 = the largest of 32768 / 24576 / 16384 / 8192 whose "
                         "steps + warm-up finish in about 150 s at the reference's ~0.65 TFLOP/s)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-also", action="store_true", help="skip the dgetrf / dgemm sub-records of the default run")
    ap.add_argument("--also-steps", type=int, default=2)
    ap.add_argument("--grid", default="", help="process grid PxQ of the default routines (default: the reference's rule, 2x4 on 8 ranks)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.routine in EXTRA_ROUTINES:
        return run_extra(args)

    import ctypes
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
    gp, gq = (int(x) for x in args.grid.lower().split("x")) if args.grid else (None, None)
    grid = sl.Grid.from_torch_distributed(gp, gq) if world > 1 else sl.Grid()
    st = torch.cuda.current_stream().cuda_stream
    env = {"torch": torch, "dist": dist, "sl": sl, "lib": lib, "check": check, "c_dbl": c_dbl, "c_ptr": c_ptr}
    warmup = max(args.warmup, 3)

    rec, mats, pair = run_factor_bench(routine, n, nb, grid, world, args.steps, warmup, env, True, local_rank)
    peak = fp64_peak_probe(lib, st)
    peak = _max_over_ranks(torch, dist, world, [peak])[0]

    def roofline_of(r, rt):
        achieved = r["trail_flops_per_step"] / (r["trail_ms_per_step"] * 1e-3) / 1e12 if r["trail_ms_per_step"] > 0 else 0.0
        traffic, alg_bytes, tsrc = measured_traffic(rt)
        return {"bound": "tensor", "kernel": f"gemm_dmma_kernel<{GEMM_VARIANT[rt]}> (trailing-update batched tile GEMM/HERK, FP64 DMMA)",
                "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak if peak else None,
                "traffic": traffic, "algorithmic_bytes": alg_bytes, "traffic_source": tsrc,
                "peak_source": "FP64 DMMA.8x8x4 probe measured live in this run (sb200_fp64_peak_probe); "
                               "MEASURED_PEAKS.json carries no FP64 figure",
                "launches_timed": r["trail_launches"], "trailing_ms_per_step": r["trail_ms_per_step"],
                "panel_stream_ms_per_step": r["panel_stream_ms_per_step"],
                "whole_step_frac_of_peak": r["value"] / (world * peak) if peak else None}

    roofline = roofline_of(rec, routine)

    # ---- e2e (potrf): public API with HOST buffers, copies inside the timed region
    e2e = None
    if not args.no_e2e and routine == "potrf":
        e2e = e2e_potrf(n, nb, grid, world, min(args.steps, 3), env, pair)
    elif not args.no_e2e:
        e2e = {"value": None, "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
               "note": "the end-to-end leg is measured on the default routine (potrf)"}
    for m in mats:
        m.close()
    del mats, pair
    torch.cuda.empty_cache()

    # ---- the metric's other two routines at the same n, timed in the same run (fewer steps)
    also = {}
    if routine == "potrf" and not args.no_also:
        for rt in ("getrf", "gemm"):
            r2, m2, _ = run_factor_bench(rt, n, nb, grid, world, max(1, min(args.also_steps, args.steps)), 1, env)
            for m in m2:
                m.close()
            del m2, _
            torch.cuda.empty_cache()
            also["d" + rt] = {"value": r2["value"], "unit": "TFLOP/s", "ms_per_step": r2["ms_per_step"],
                              "device_ms_per_step": r2["device_ms_per_step"], "steps": r2["steps"], "warmup": r2["warmup"],
                              "frac_of_aggregate_peak": r2["value"] / (world * peak) if peak else None,
                              "roofline": roofline_of(r2, rt), "check": r2["check"], "gpu_launches": r2["gpu_launches"],
                              "config": workload_config(rt, n, nb, grid.p, grid.q, world)}

    # ---- CPU baseline: reference HostTask on this box's cores, bounded sample (rank 0, N = 1)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            threads = os.cpu_count() or 1
            # one run of about 10 s of host work on 16 threads: dpotrf 11.7 TFLOP at 1.1 TFLOP/s; dgetrf 2.9 at 0.45; dgemm 8.8 at 0.8
            ref_n = args.ref_n or (32768 if routine == "potrf" else 16384)
            secs, kind = cpu_reference_run(routine, ref_n, nb, threads)
            cpu = {"value": flops(routine, ref_n) / secs / 1e12, "unit": "TFLOP/s", "cores": threads, "kind": kind,
                   "sample": f"d{routine} n={ref_n} nb={nb} Target::HostTask (the workload's generator and tile size at a "
                             f"bounded n), one run, {secs:.2f} s"}
        except Exception as ex:   # noqa: BLE001
            cpu = {"value": None, "unit": "TFLOP/s", "cores": os.cpu_count(), "kind": "reference", "sample": f"failed: {ex}"}

    if rank == 0:
        cfg = workload_config(routine, n, nb, grid.p, grid.q, world)
        line = {
            "metric": f"d{routine} TFLOP/s", "value": rec["value"], "unit": "TFLOP/s", "n_gpus": world,
            "steps": args.steps, "warmup": warmup, "ms_per_step": rec["ms_per_step"],
            "device_ms_per_step": rec["device_ms_per_step"], "restore_ms_per_step": rec["restore_ms_per_step"],
            "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64",
            "data": "synthetic (reference matgen: Philox-2x64 rand_dominant/rand, seed 42, generated on device)",
            "config": cfg, "switches": _switches(), "frac_of_aggregate_peak": rec["value"] / (world * peak) if peak else None,
            "clocks": rec["clocks"], "roofline": roofline, "check": rec["check"], "also": also,
            "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": rec["gpu_launches"],
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
