"""Compare the SASS of two builds of libslate_b200.so kernel by kernel (no GPU needed).

    python scratch/sass_compare.py OLD.so NEW.so

Every kernel is identified by its demangled name; its body is the sequence of instruction lines of `cuobjdump -sass`
(opcode + operands, addresses and encodings stripped).  Prints how many kernels are bitwise identical, which differ, which
are gone and which are new -- the check behind profiles/*_sass_unchanged_check.txt: a change of host code or of one
kernel must leave every measured kernel as it was."""
import re
import subprocess
import sys


def kernels(path):
    txt = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
    out, name, body = {}, None, []
    for line in txt.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            if name is not None:
                out[name] = body
            name, body = m.group(1), []
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(.*?);", line)
        if m and name is not None:
            body.append(re.sub(r"\s+", " ", m.group(1)))
    if name is not None:
        out[name] = body
    return out


def demangle(names):
    r = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True)
    return dict(zip(names, r.stdout.splitlines())) if r.returncode == 0 else {n: n for n in names}


def by_demangled(ks):
    """anonymous-namespace kernels carry a per-translation-unit hash of the source PATH in their mangled names: key on
    the demangled name so that builds from two directories compare"""
    dm = demangle(list(ks))
    return {dm[k]: v for k, v in ks.items()}


def main():
    old, new = by_demangled(kernels(sys.argv[1])), by_demangled(kernels(sys.argv[2]))
    same = [k for k in old if k in new and old[k] == new[k]]
    differ = [k for k in old if k in new and old[k] != new[k]]
    gone = [k for k in old if k not in new]
    added = [k for k in new if k not in old]
    dm = {k: k for k in differ + gone + added}
    print(f"{len(old)} kernels before; {len(same)} bitwise identical SASS; {len(differ)} differ; {len(gone)} gone; {len(added)} new")
    for title, ks in (("differ", differ), ("gone", gone), ("new", added)):
        for k in ks:
            print(f"  {title}: {dm[k]}  ({len(new.get(k, old.get(k)))} instructions)")
    return 1 if differ or gone else 0


if __name__ == "__main__":
    sys.exit(main())
