#!/usr/bin/env python
"""Turn gpurun_out/ ncu artefacts into the tracked summaries under profiles/.
usage: python scratch/summarize_profiles.py r01 launches_potrf.csv prof_gemm_dmma.ncu-rep"""
import collections, csv, json, re, subprocess, sys

tag, launches, rep = sys.argv[1], sys.argv[2], sys.argv[3]
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "lts__t_sector_hit_rate.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_elapsed.max",
        "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
        "smsp__warp_issue_stalled_barrier_per_warp_active.pct",
        "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct",
        "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct",
        "smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct"]

if launches != "-":
    lines = [l for l in open(f"gpurun_out/{launches}") if not l.startswith("==")]
    tot, cnt = collections.defaultdict(float), collections.Counter()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", row["Kernel Name"])
        name = re.sub(r"^void ", "", name)
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}[row["Metric Unit"]]
        tot[name] += v; cnt[name] += 1
    T = sum(tot.values())
    out = f"profiles/{tag}_{launches.replace('.csv', '')}_summary.txt"
    with open(out, "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)\n")
        f.write(f"# source: gpurun_out/{launches}; {sum(cnt.values())} launches, {T / 1e6:.2f} ms summed device time\n")
        f.write(f"{'kernel':90s} {'launches':>9s} {'ms':>10s} {'share':>7s} {'us/launch':>10s}\n")
        for k, v in sorted(tot.items(), key=lambda x: -x[1]):
            f.write(f"{k[:90]:90s} {cnt[k]:9d} {v / 1e6:10.2f} {100 * v / T:6.2f}% {v / 1e3 / cnt[k]:10.1f}\n")
    print(open(out).read())

if rep != "-":
    raw = subprocess.run(["ncu", "-i", f"gpurun_out/{rep}", "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    out = f"profiles/{tag}_{rep.replace('.ncu-rep', '')}_ncu.txt"
    with open(out, "w") as f:
        f.write(f"# ncu --set full --clock-control none --import-source on; source gpurun_out/{rep}\n")
        for row in rows[2:]:
            d = dict(zip(hdr, row))
            f.write(f"\n## {d['Kernel Name']}  grid {d.get('Grid Size')} block {d.get('Block Size')}\n")
            for i, h in enumerate(hdr):
                if h in KEYS:
                    f.write(f"{h:90s} {row[i]:>16s} {units[i]}\n")
            last = d
    print(open(out).read())
