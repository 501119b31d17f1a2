"""CPU emulation of the PROTOCOL of getrf_base_ll_kernel (slate_b200/csrc/getrf.cu): G "CTAs" run as Python threads,
each owning a contiguous range of panel rows, and exchange their per-column candidates, the diagonal row and the
winner's row through tagged slots {data, generation} double-buffered by column parity -- no barrier, exactly as the
kernel does.  Random delays shuffle the interleaving.  Checked: no deadlock, no torn / stale read (every poll accepts
a slot only when its tag equals the column's tag), pivots and factors equal to a plain partial-pivoting LU with the
reference's tie rule (first maximum; the diagonal wins ties), several launches in a row (growing tags), a row-less
interchange CTA included.
usage: python scratch/emulate_panel_ll.py"""
import random
import threading
import time

import numpy as np

PW = 32


class Slots:
    """rec[par][cta] = (val, row, tag); cand[par][cta][c] = (value, tag); diag[par][c] = (value, tag).  A slot is
    replaced by one reference assignment of a tuple -- the analogue of the kernel's single-copy-atomic 8-byte word."""

    def __init__(self, G):
        self.rec = [[(0.0, 0, 0)] * G for _ in range(2)]
        self.cand = [[[(0.0, 0)] * PW for _ in range(G)] for _ in range(2)]
        self.diag = [[(0.0, 0)] * PW for _ in range(2)]


def cta(b, G, blk, r_begin, m_p, c0, w, slots, gen_base, out_piv, jitter, deadline):
    """blk: this CTA's rows (nr x w, numpy, modified in place)."""
    nr = blk.shape[0]
    r_end = r_begin + nr
    rnd = random.Random(1000 * gen_base + b)

    def wait(get, gen):
        while True:
            v = get()
            if v[-1] == gen:
                return v
            if time.time() > deadline:
                raise RuntimeError("deadlock / lost update in the emulated protocol")
            time.sleep(0)

    for j in range(w):
        d = c0 + j
        par = j & 1
        gen = gen_base + j + 1
        if jitter:
            time.sleep(rnd.random() * jitter)
        # local candidate: first maximum below the diagonal
        best, brow = -1.0, 2 ** 31 - 1
        for lr in range(nr):
            r = r_begin + lr
            if r > d:
                v = abs(blk[lr, j])
                if v > best:
                    best, brow = v, r
        slots.rec[par][b] = (best, brow, gen)
        if brow != 2 ** 31 - 1:
            for c in range(w):
                slots.cand[par][b][c] = (blk[brow - r_begin, c], gen)
        if r_begin <= d < r_end:
            for c in range(w):
                slots.diag[par][c] = (blk[d - r_begin, c], gen)
        # gather
        bv, br, bw = -1.0, 2 ** 31 - 1, -1
        for c in range(G):
            v, rr, _ = wait(lambda c=c: slots.rec[par][c], gen)
            if v > bv or (v == bv and rr < br):
                bv, br, bw = v, rr, c
        drow = [wait(lambda c=c: slots.diag[par][c], gen)[0] for c in range(w)]
        if bv > abs(drow[j]):
            p, sw = br, bw
        else:
            p, sw = d, -1
        prow = drow if p == d else [wait(lambda c=c: slots.cand[par][sw][c], gen)[0] for c in range(w)]
        if p != d:
            if r_begin <= p < r_end:
                blk[p - r_begin, :] = drow
            if r_begin <= d < r_end:
                blk[d - r_begin, :] = prow
        if b == 0:
            out_piv[d] = p
        pv = prow[j]
        if pv != 0.0:
            for lr in range(nr):
                r = r_begin + lr
                if r > d:
                    l = blk[lr, j] / pv
                    blk[lr, j] = l
                    blk[lr, j + 1:] -= l * np.asarray(prow[j + 1:])


def run_launch(A, c0, w, rows_per, gen_base, slots, piv, jitter, extra_cta):
    m_p = A.shape[0]
    G = -(-(m_p - c0) // rows_per) + (1 if extra_cta else 0)
    blks, threads, errs = [], [], []
    deadline = time.time() + 60
    for b in range(G):
        r_begin = c0 + b * rows_per
        r_end = min(r_begin + rows_per, m_p)
        blk = A[r_begin:max(r_end, r_begin), c0:c0 + w].copy()
        blks.append((r_begin, blk))

        def body(b=b, blk=blk, r_begin=r_begin):
            try:
                cta(b, G, blk, r_begin, m_p, c0, w, slots, gen_base, piv, jitter, deadline)
            except Exception as ex:  # noqa: BLE001
                errs.append(ex)
        threads.append(threading.Thread(target=body))
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errs, errs
    for r_begin, blk in blks:
        A[r_begin:r_begin + blk.shape[0], c0:c0 + w] = blk
    return G


def reference_block(A, c0, w):
    """plain partial pivoting on columns [c0, c0+w) over rows [c0, m): swaps restricted to these columns, as the kernel"""
    m = A.shape[0]
    piv = {}
    for j in range(w):
        d = c0 + j
        col = np.abs(A[d + 1:, c0 + j])
        p = d
        if col.size and col.max() > abs(A[d, c0 + j]):
            p = d + 1 + int(np.argmax(col))            # first maximum
        piv[d] = p
        if p != d:
            A[[d, p], c0:c0 + w] = A[[p, d], c0:c0 + w]
        pv = A[d, c0 + j]
        if pv != 0.0:
            A[d + 1:, c0 + j] /= pv
            A[d + 1:, c0 + j + 1:c0 + w] -= np.outer(A[d + 1:, c0 + j], A[d, c0 + j + 1:c0 + w])
    return piv


def main():
    rng = np.random.default_rng(0)
    for m_p, rows_per, w, kind, extra in ((96, 17, 32, "rand", True), (70, 70, 16, "rand", False), (120, 13, 32, "ties", True),
                                          (64, 9, 32, "zero", False)):
        A = rng.random((m_p, 2 * w))
        if kind == "ties":
            A = np.sign(A - 0.5)
        if kind == "zero":
            A[:, 5] = 0.0
        Aref = A.copy()
        slots = Slots(-(-m_p // rows_per) + 1)
        gen_base = 0
        piv = {}
        for launch, c0 in enumerate((0, w)):                 # two launches in a row: tags keep growing, slots are reused
            G = run_launch(A, c0, w, rows_per, gen_base, slots, piv, jitter=2e-4 if launch == 0 else 0.0, extra_cta=extra)
            gen_base += PW
            pref = reference_block(Aref, c0, w)
            assert all(piv[d] == pref[d] for d in pref), (kind, launch)
            assert np.array_equal(A[:, c0:c0 + w], Aref[:, c0:c0 + w]) or \
                np.abs(A[:, c0:c0 + w] - Aref[:, c0:c0 + w]).max() < 1e-13, (kind, launch)
            # bring the next block of both copies to the same state (the drivers' trsm + gemm between launches are
            # not part of this emulation: the second launch simply factors the untouched next block of columns)
            Aref[:, w:] = A[:, w:] if launch == 0 else Aref[:, w:]
        print(f"panel LL protocol: m_p={m_p} rows/CTA={rows_per} w={w} {kind}: {G} CTAs, pivots and factors OK")


if __name__ == "__main__":
    main()
