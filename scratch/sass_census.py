#!/usr/bin/env python
"""Per-kernel opcode census of the shipped library (no GPU needed): cuobjdump -sass slate_b200/lib/libslate_b200.so.
Writes profiles/r02_sass_opcode_census.txt.  The opcodes listed per kernel are the ones that prove the data path
(tensor-core MMAs, TMEM accesses, TMA bulk copies, mbarrier ops, ...) plus the arithmetic and memory instructions."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "slate_b200", "lib", "libslate_b200.so")
KEEP = ("DMMA", "HMMA", "UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UBLKCP", "UTMALDG", "UTCBAR", "UTCATOMSWS", "SYNCS", "LDGSTS", "DFMA", "DMUL",
        "DADD", "FFMA", "LDG", "STG", "LDS", "STS", "MUFU", "BAR", "MEMBAR", "REDUX", "SHFL", "WARPSYNC", "USETMAXREG", "ATOM", "RED", "CCTL")
HEAD = """# cuobjdump -sass slate_b200/lib/libslate_b200.so (sm_100a), per-kernel opcode census of the instructions that prove the data path:
# DMMA = FP64 tensor-core MMA (8x8x4), UTCHMMA = tcgen05.mma (kind::tf32/f16), LDTM/STTM = tcgen05.ld/st (TMEM), UBLKCP = cp.async.bulk (1-D TMA bulk copy),
# UTMALDG = tensor-map TMA (none: tiles are addressed through pointer arrays, one 1-D bulk copy per tile row segment), SYNCS = mbarrier ops, LDGSTS = cp.async
# (written by scratch/sass_census.py)
"""


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur:
            kernels[cur][m.group(1).split(".")[0]] += 1
            kernels[cur]["_n"] += 1
    names = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    lines = [HEAD]
    for (mangled, cnt), name in sorted(zip(kernels.items(), names), key=lambda x: x[1]):
        name = re.sub(r"\(.*$", "", name.replace("(anonymous namespace)::", ""))
        ops = sorted(((k, v) for k, v in cnt.items() if k in KEEP), key=lambda kv: -kv[1])
        lines.append(f"{name}\n    instructions {cnt['_n']}: " + ", ".join(f"{k} {v}" for k, v in ops) + "\n")
    path = os.path.join(ROOT, "profiles", "r02_sass_opcode_census.txt")
    open(path, "w").write("\n".join(lines))
    print(path, len(kernels), "kernels")


if __name__ == "__main__":
    sys.exit(main())
