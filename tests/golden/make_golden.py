#!/usr/bin/env python
"""Generates the committed golden fixtures in tests/golden/ from the UNMODIFIED reference.

Runs oracle/_ref/ref_dump (built from /root/reference by oracle/build_ref.sh: the reference's own
matgen + Target::HostTask drivers, OpenMP + OpenBLAS) for small seeded problems and stores inputs
are NOT stored (they are regenerated bit-exactly from the seed by the Philox generator; one
generator fixture pins that), only outputs:

  gen_{s,d,c,z}.npz      Philox matrices (rand and rand_dominant), bit-exact pin of matgen
  potrf_d.npz            L = chol(A) lower, n=384 nb=128, rand_dominant seed 42
  getrf_d.npz            LU and slate::Pivots (tileIndex, elementOffset), n=384 nb=128 ib=16, rand seed 42
  getrf_d_ragged.npz     same, n=300 nb=128 (ragged last tile)
  gemm_d.npz, gemm_z.npz C = alpha A B + beta C, tester alpha/beta, n=256 nb=64
  herk_d.npz, herk_z.npz C = alpha A A^H + beta C lower, n=256 k=128 nb=64
  trsm_d.npz             Left/Lower/NoTrans/NonUnit solve, m=256 n=128 nb=64
  norms_d.npz            max/one/inf/fro of rand 200x136
  gesv_mixed_d.npz       solution + iteration count, n=256 nb=64
  posv_mixed_d.npz       the same for the Cholesky mixed solver (rand_dominant), n=256 nb=64
  posv_d.npz, posv_z.npz chol_factor + chol_solve_using_factor solution, n=300 nb=128 / n=192 nb=64 nrhs=70
  gesv_d.npz             lu_factor + lu_solve_using_factor solution, n=300 nb=128
  hemm_z.npz             C = alpha A B + beta C, A Hermitian lower, n=192 nb=64 nrhs=70
  potrf_z.npz            complex Cholesky factor, n=192 nb=64
  trmm_d.npz             B = alpha A B, A lower triangular (rand), Left/NoTrans/NonUnit, m=200 n=70 nb=64
  symm_z.npz             C = alpha A B + beta C, A complex-symmetric lower, n=192 nb=64 nrhs=70
  syrk_z.npz, syr2k_z.npz   complex-symmetric rank-k / rank-2k updates (no conjugation), n=200 k=100 nb=64
  getrf_nopiv_d.npz      LU without pivoting, rand_dominant, n=300 nb=128 (ragged)
  her2k_d.npz, her2k_z.npz  C = alpha A B^H + conj(alpha) B A^H + beta C lower, n=200 k=100 nb=64 (ragged tiles)
  getrf_z.npz, getrf_c.npz  complex LU and pivots (cabs1 rule), n=192 / 200 (ragged) nb=64; gesv_z.npz its solve, n=200 nrhs=70;
                            gesv_mixed_z.npz, posv_mixed_z.npz complex mixed solvers (solution + iteration count), n=256 nb=64;
                            getrf_tntpiv_z.npz (CALU, n=192), getrf_nopiv_z.npz (rand_dominant, n=200)
  herk_{z_conj,d_trans}.npz, her2k_z_conj.npz, syrk_z_trans.npz, syr2k_z_trans.npz   rank-k / rank-2k updates with A (and B) stored
                            k x n and handed over as (conjugate-)transposed views, n=200 k=100 nb=64
  gesv_d_trans.npz, gesv_z_conj.npz   lu_factor, then lu_solve_using_factor with the (conjugate-)transposed view: op(A) X = B
  gemm_{d_tn,d_nt,z_cn,z_tc,z_nc}.npz   slate::multiply with (conjugate-)transposed views of A / B, m=150 n=200 k=100 nb=64
  {trmm,trsm}_{z_left_conj,d_left_trans,d_right,z_right_trans,z_right_conj}.npz, hemm_{z,d}_right.npz, symm_z_right.npz
                            the other side / op variants of trmm / hemm / symm (lower storage), nb=64
  grid_*.npz                the reference ON PROCESS GRIDS (oracle/_ref/ref_dump_mp under oracle/mprun.py): getrf_tntpiv on 2x1 / 3x1 /
                            4x1 / 2x4 ranks (the tournament proper), getrf on 2x2 / 3x2, potrf on 2x2; nb=64
  getrf_tntpiv_d{,_ragged,_tall}.npz   LU with tournament pivoting (MethodLU::CALU), 384x384 / 300x300 / 512x256, nb=128

Usage (in the build container, where /root/reference exists):  python tests/golden/make_golden.py
"""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
EXE = os.path.join(ROOT, "oracle", "_ref", "ref_dump")
OUT = os.path.dirname(os.path.abspath(__file__))
DT = {"s": np.float32, "d": np.float64, "c": np.complex64, "z": np.complex128}


def run(routine, t, n, nb, seeds=(42, 43, 44), **kv):
    tmp = tempfile.mkdtemp()
    prefix = os.path.join(tmp, "x")
    cmd = [EXE, routine, t, str(n), str(nb), *map(str, seeds), prefix] + [f"{k}={v}" for k, v in kv.items()]
    env = dict(os.environ, OPENBLAS_NUM_THREADS="1", OMP_NUM_THREADS="4")
    r = subprocess.run(cmd, capture_output=True, text=True, env=env, check=True)
    meta = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    files = {}
    for f in os.listdir(tmp):
        key = f.split(".")[1]
        dtype = np.int64 if key == "piv" else (np.float64 if (routine == "norms" and key == "out") else DT[t])
        files[key] = np.fromfile(os.path.join(tmp, f), dtype=dtype)
    return files, meta


MPRUN = os.path.join(ROOT, "oracle", "mprun.py")
EXE_MP = os.path.join(ROOT, "oracle", "_ref", "ref_dump_mp")


def tile_owner(i, j, p, q):
    return (i % p) + (j % q) * p            # GridOrder::Col (include/slate/func.hh:96-104)


def run_mp(routine, t, n, nb, p, q, seeds=(42, 43, 44), m=None, rows=None, **kv):
    """The unmodified reference on a p x q process grid: oracle/_ref/ref_dump_mp (built by oracle/build_ref_mp.sh against the
    multi-process MPI replacement oracle/mpi_mp) under oracle/mprun.py.  Every rank writes the tiles it owns; they are
    assembled here with the reference's tile map.  Returns ({name: array}, meta)."""
    tmp = tempfile.mkdtemp()
    prefix = os.path.join(tmp, "x")
    cmd = [sys.executable, MPRUN, "-n", str(p * q), "--timeout", "150", EXE_MP, routine, t, str(n), str(nb), *map(str, seeds), prefix,
           f"p={p}", f"q={q}"] + ([f"m={m}"] if m is not None else []) + [f"{k}={v}" for k, v in kv.items()]
    m = n if m is None else m
    env = dict(os.environ, OPENBLAS_NUM_THREADS="1", OMP_NUM_THREADS="2")
    r = subprocess.run(cmd, capture_output=True, text=True, env=env, check=True, timeout=200)
    meta = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    out = {}
    names = {f.split(".")[2] for f in os.listdir(tmp)}
    for name in names:
        if name == "piv":
            out[name] = np.fromfile(f"{prefix}.r0.piv.bin", dtype=np.int64)
            continue
        nrows = rows if rows is not None else (m if routine in ("getrf", "gemm") else n)
        parts = [np.fromfile(f"{prefix}.r{k}.{name}.bin", dtype=DT[t]).reshape(nrows, -1, order="F") for k in range(p * q)]
        full = np.zeros_like(parts[0])
        for j in range(-(-full.shape[1] // nb)):
            for i in range(-(-nrows // nb)):
                full[i * nb:(i + 1) * nb, j * nb:(j + 1) * nb] = parts[tile_owner(i, j, p, q)][i * nb:(i + 1) * nb, j * nb:(j + 1) * nb]
        out[name] = full
    return out, meta


def grid_fixtures():
    """Outputs of the reference itself on process grids (tournament pivoting needs >= 2 ranks in a process column to be
    more than partial pivoting; the cross-rank pivot rule of getrf and the Cholesky / SUMMA data flow likewise)."""
    if not os.path.exists(EXE_MP):
        sys.exit("oracle/_ref/ref_dump_mp missing: run oracle/build_ref_mp.sh first")
    for name, p, q, m, n, full in (("grid_getrf_tntpiv_d_2x1", 2, 1, 384, 384, True), ("grid_getrf_tntpiv_d_3x1_ragged", 3, 1, 300, 300, True),
                                   ("grid_getrf_tntpiv_d_4x1_tall", 4, 1, 448, 256, True), ("grid_getrf_tntpiv_d_2x4", 2, 4, 512, 512, False)):
        f, meta = run_mp("getrf", "d", n, 64, p, q, m=m, ib=16, pt=1, method="calu")
        extra = {"out": f["out"]} if full else {"diag_u": np.diag(f["out"]).copy(), "abs_sum": np.abs(f["out"]).sum()}
        np.savez_compressed(os.path.join(OUT, f"{name}.npz"), piv=f["piv"].reshape(-1, 2), info=meta["info"], **extra)
    f, meta = run_mp("getrf", "d", 384, 64, 2, 2, ib=16, pt=1)
    np.savez_compressed(os.path.join(OUT, "grid_getrf_d_2x2.npz"), out=f["out"], piv=f["piv"].reshape(-1, 2), info=meta["info"])
    f, meta = run_mp("getrf", "d", 300, 64, 3, 2, seeds=(7, 43, 44), m=500, ib=16, pt=1)
    np.savez_compressed(os.path.join(OUT, "grid_getrf_d_3x2_tall.npz"), piv=f["piv"].reshape(-1, 2), info=meta["info"],
                        diag_u=np.diag(f["out"]).copy(), abs_sum=np.abs(f["out"]).sum())
    f, meta = run_mp("potrf", "d", 384, 64, 2, 2)
    np.savez_compressed(os.path.join(OUT, "grid_potrf_d_2x2.npz"), out=np.tril(f["out"]), info=meta["info"])


def tntpiv_fixtures():
    for name, m, n in (("getrf_tntpiv_d", 384, 384), ("getrf_tntpiv_d_ragged", 300, 300), ("getrf_tntpiv_d_tall", 512, 256)):
        f, meta = run("getrf", "d", n, 128, ib=16, pt=1, method="calu", m=m)
        np.savez_compressed(os.path.join(OUT, f"{name}.npz"), out=f["out"].reshape(m, n, order="F"),
                            piv=f["piv"].reshape(-1, 2), info=meta["info"])


def complex_lu_fixtures():
    f, meta = run("getrf", "z", 192, 64, ib=16, pt=1)
    np.savez_compressed(os.path.join(OUT, "getrf_z.npz"), out=f["out"].reshape(192, 192, order="F"),
                        piv=f["piv"].reshape(-1, 2), info=meta["info"])
    f, meta = run("getrf", "c", 200, 64, ib=16, pt=1)
    np.savez_compressed(os.path.join(OUT, "getrf_c.npz"), out=f["out"].reshape(200, 200, order="F"),
                        piv=f["piv"].reshape(-1, 2), info=meta["info"])
    f, meta = run("gesv", "z", 200, 64, ib=16, pt=1, nrhs=70)
    np.savez_compressed(os.path.join(OUT, "gesv_z.npz"), out=f["out"].reshape(200, 70, order="F"), info=meta["info"])
    f, meta = run("getrf", "z", 192, 64, ib=16, pt=1, method="calu")
    np.savez_compressed(os.path.join(OUT, "getrf_tntpiv_z.npz"), out=f["out"].reshape(192, 192, order="F"),
                        piv=f["piv"].reshape(-1, 2), info=meta["info"])
    f, meta = run("getrf_nopiv", "z", 200, 64)
    np.savez_compressed(os.path.join(OUT, "getrf_nopiv_z.npz"), out=f["out"].reshape(200, 200, order="F"), info=meta["info"])
    # complex mixed-precision solvers <complex<double>, complex<float>> (src/gesv_mixed.cc:303-316, src/posv_mixed.cc)
    for r in ("gesv_mixed", "posv_mixed"):
        f, meta = run(r, "z", 256, 64, ib=16, pt=1)
        np.savez_compressed(os.path.join(OUT, f"{r}_z.npz"), out=f["out"].reshape(256, 10, order="F"),
                            iters=meta["iters"], info=meta["info"])


BLAS3_VARIANTS = [
    # name, routine, type, kv of ref_dump: trmm is m x n (n positional), hemm / symm n x n with nrhs rows or columns
    ("trmm_z_left_conj",   "trmm", "z", dict(n=70,  m=200, op="c")),
    ("trmm_d_left_trans",  "trmm", "d", dict(n=70,  m=200, op="t", diag="u")),
    ("trmm_d_right",       "trmm", "d", dict(n=200, m=70)),
    ("trmm_z_right_trans", "trmm", "z", dict(n=200, m=70, op="t")),
    ("trmm_z_right_conj",  "trmm", "z", dict(n=200, m=70, op="c", diag="u")),
    ("gemm_d_tn",          "gemm", "d", dict(n=200, m=150, k=100, opa="t")),
    ("gemm_d_nt",          "gemm", "d", dict(n=200, m=150, k=100, opb="t")),
    ("gemm_z_cn",          "gemm", "z", dict(n=200, m=150, k=100, opa="c")),
    ("gemm_z_tc",          "gemm", "z", dict(n=200, m=150, k=100, opa="t", opb="c")),
    ("gemm_z_nc",          "gemm", "z", dict(n=200, m=150, k=100, opb="c")),
    ("herk_z_conj",        "herk",  "z", dict(n=200, k=100, trans="c")),
    ("herk_d_trans",       "herk",  "d", dict(n=200, k=100, trans="c")),
    ("her2k_z_conj",       "her2k", "z", dict(n=200, k=100, trans="c")),
    ("syrk_z_trans",       "syrk",  "z", dict(n=200, k=100, trans="t")),
    ("syr2k_z_trans",      "syr2k", "z", dict(n=200, k=100, trans="t")),
    ("gesv_d_trans",       "gesv", "d", dict(n=300, nrhs=70, trans="t", ib=16, pt=1)),
    ("gesv_z_conj",        "gesv", "z", dict(n=200, nrhs=70, trans="c", ib=16, pt=1)),
    ("trsm_z_left_conj",   "trsm", "z", dict(n=70,  m=200, op="c")),
    ("trsm_d_left_trans",  "trsm", "d", dict(n=70,  m=200, op="t", diag="u")),
    ("trsm_d_right",       "trsm", "d", dict(n=200, m=70)),
    ("trsm_z_right_trans", "trsm", "z", dict(n=200, m=70, op="t", diag="u")),
    ("trsm_z_right_conj",  "trsm", "z", dict(n=200, m=70, op="c")),
    ("hemm_z_right",       "hemm", "z", dict(n=192, nrhs=70, side="r")),
    ("hemm_d_right",       "hemm", "d", dict(n=200, nrhs=70, side="r")),
    ("symm_z_right",       "symm", "z", dict(n=192, nrhs=70, side="r")),
]


def blas3_variant_fixtures():
    """SURVEY section 8(f) items 1 and 3, the other side / op variants: slate::trmm and slate::trsm with Side::Right and with
    transposed / conjugate-transposed views of a lower-triangular A (trsm: rand_dominant), slate::hemm / slate::symm with
    Side::Right (nb = 64)."""
    for name, routine, t, kv in BLAS3_VARIANTS:
        kv = dict(kv)
        n = kv.pop("n")
        if routine in ("trmm", "trsm") and name.split("_")[2] == "right":
            kv["side"] = "r"
        f, _ = run(routine, t, n, 64, **kv)
        if routine == "gesv":
            shape = (n, kv["nrhs"])
        elif routine in ("herk", "her2k", "syrk", "syr2k"):
            shape = (n, n)
        elif routine in ("trmm", "trsm", "gemm"):
            shape = (kv["m"], n)
        else:
            shape = (kv["nrhs"], n)            # Side::Right: B and C are nrhs x n
        np.savez_compressed(os.path.join(OUT, f"{name}.npz"), out=f["out"].reshape(shape, order="F"))


def main():
    if not os.path.exists(EXE):
        sys.exit("oracle/_ref/ref_dump missing: run oracle/build_ref.sh first")
    for t in "sdcz":
        a, _ = run("gen", t, 96, 32, seeds=(5, 6, 7), kind="rand", m=80)
        b, _ = run("gen", t, 96, 32, seeds=(7, 6, 7), kind="rand_dominant")
        np.savez_compressed(os.path.join(OUT, f"gen_{t}.npz"), rand=a["A"].reshape(80, 96, order="F"),
                            rand_dominant=b["A"].reshape(96, 96, order="F"))
    f, meta = run("potrf", "d", 384, 128)
    np.savez_compressed(os.path.join(OUT, "potrf_d.npz"), out=f["out"].reshape(384, 384, order="F"), info=meta["info"])
    for name, n in (("getrf_d", 384), ("getrf_d_ragged", 300)):
        f, meta = run("getrf", "d", n, 128, ib=16, pt=1)
        np.savez_compressed(os.path.join(OUT, f"{name}.npz"), out=f["out"].reshape(n, n, order="F"),
                            piv=f["piv"].reshape(-1, 2), info=meta["info"])
    for t in "dz":
        f, _ = run("gemm", t, 256, 64)
        np.savez_compressed(os.path.join(OUT, f"gemm_{t}.npz"), out=f["out"].reshape(256, 256, order="F"))
        f, _ = run("herk", t, 256, 64, k=128)
        np.savez_compressed(os.path.join(OUT, f"herk_{t}.npz"), out=f["out"].reshape(256, 256, order="F"))
    f, _ = run("trsm", "d", 128, 64, m=256)
    np.savez_compressed(os.path.join(OUT, "trsm_d.npz"), out=f["out"].reshape(256, 128, order="F"))
    f, _ = run("norms", "d", 136, 64, m=200)
    np.savez_compressed(os.path.join(OUT, "norms_d.npz"), out=f["out"])
    f, meta = run("gesv_mixed", "d", 256, 64)
    np.savez_compressed(os.path.join(OUT, "gesv_mixed_d.npz"), out=f["out"].reshape(256, 10, order="F"),
                        iters=meta["iters"], info=meta["info"])
    # solve path and Hermitian mixed solver (round-1 widening)
    f, meta = run("posv_mixed", "d", 256, 64)
    np.savez_compressed(os.path.join(OUT, "posv_mixed_d.npz"), out=f["out"].reshape(256, 10, order="F"),
                        iters=meta["iters"], info=meta["info"])
    f, meta = run("posv", "d", 300, 128)
    np.savez_compressed(os.path.join(OUT, "posv_d.npz"), out=f["out"].reshape(300, 10, order="F"), info=meta["info"])
    f, meta = run("posv", "z", 192, 64, nrhs=70)
    np.savez_compressed(os.path.join(OUT, "posv_z.npz"), out=f["out"].reshape(192, 70, order="F"), info=meta["info"])
    f, meta = run("gesv", "d", 300, 128, ib=16, pt=1)
    np.savez_compressed(os.path.join(OUT, "gesv_d.npz"), out=f["out"].reshape(300, 10, order="F"), info=meta["info"])
    f, _ = run("hemm", "z", 192, 64, nrhs=70)
    np.savez_compressed(os.path.join(OUT, "hemm_z.npz"), out=f["out"].reshape(192, 70, order="F"))
    f, meta = run("potrf", "z", 192, 64)
    np.savez_compressed(os.path.join(OUT, "potrf_z.npz"), out=f["out"].reshape(192, 192, order="F"), info=meta["info"])
    # section 8(f) item 3 widening: her2k / syr2k
    for t in "dz":
        f, _ = run("her2k", t, 200, 64, k=100)
        np.savez_compressed(os.path.join(OUT, f"her2k_{t}.npz"), out=f["out"].reshape(200, 200, order="F"))
    f, _ = run("trmm", "d", 70, 64, m=200)
    np.savez_compressed(os.path.join(OUT, "trmm_d.npz"), out=f["out"].reshape(200, 70, order="F"))
    f, _ = run("symm", "z", 192, 64, nrhs=70)
    np.savez_compressed(os.path.join(OUT, "symm_z.npz"), out=f["out"].reshape(192, 70, order="F"))
    f, _ = run("syrk", "z", 200, 64, k=100)
    np.savez_compressed(os.path.join(OUT, "syrk_z.npz"), out=f["out"].reshape(200, 200, order="F"))
    f, _ = run("syr2k", "z", 200, 64, k=100)
    np.savez_compressed(os.path.join(OUT, "syr2k_z.npz"), out=f["out"].reshape(200, 200, order="F"))
    blas3_variant_fixtures()
    # section 8(f) item 2 widening: LU without pivoting
    f, meta = run("getrf_nopiv", "d", 300, 128)
    np.savez_compressed(os.path.join(OUT, "getrf_nopiv_d.npz"), out=f["out"].reshape(300, 300, order="F"), info=meta["info"])
    # section 8(f) item 2 widening: LU with tournament pivoting (one rank: the serial MPI stub)
    tntpiv_fixtures()
    # complex LU (cabs1 pivot rule, src/internal/Tile_getrf.hh:210-237) and its solve
    complex_lu_fixtures()
    # the reference on process grids (multi-process MPI replacement, oracle/mpi_mp)
    grid_fixtures()
    print("golden fixtures written to", OUT)

if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "tntpiv":
        tntpiv_fixtures()
    elif len(sys.argv) > 1 and sys.argv[1] == "complex_lu":
        complex_lu_fixtures()
    elif len(sys.argv) > 1 and sys.argv[1] == "grid":
        grid_fixtures()
    elif len(sys.argv) > 1 and sys.argv[1] == "blas3_variants":
        blas3_variant_fixtures()
    else:
        main()
