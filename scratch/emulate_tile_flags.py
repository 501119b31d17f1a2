"""CPU emulation of the FLAG PROTOCOL of potrf_tile_fused_kernel (slate_b200/csrc/potrf_tile_fused.cu): one Python thread
per CTA (64-row block), started in random order with random delays, synchronised only through rowcnt[] / diagf[]
exactly as the kernel (wait rowcnt[b] >= b before S_b, wait diagf[b] before the solve, publish rowcnt[r] = b + 1 and
diagf[r]); FAILED propagation on a non-positive-definite block.  Checked: no deadlock, L equal to LAPACK's, info of the
first failing minor, later CTAs leave without touching their flags' waiters.
usage: python scratch/emulate_tile_flags.py"""
import random
import threading
import time

import numpy as np

FB = 8            # block size of the model (the protocol does not depend on it)
FAILED = 1 << 30


def cta(r, nblk, A, W, rowcnt, diagf, info, deadline, rnd):
    def wait(arr, idx, target):
        while arr[idx] < target:
            if time.time() > deadline:
                raise RuntimeError("deadlock in the emulated flag protocol")
            time.sleep(0)
        return arr[idx]

    def bail():
        rowcnt[r] = FAILED; diagf[r] = FAILED

    blk = lambda i, j: A[i * FB:(i + 1) * FB, j * FB:(j + 1) * FB]
    time.sleep(rnd.random() * 2e-3)
    if info[0] != 0:
        bail(); return
    accD = np.zeros((FB, FB))
    for b in range(r):
        if b > 0 and wait(rowcnt, b, b) & FAILED:
            bail(); return
        acc = np.zeros((FB, FB))
        for c in range(b):
            acc += blk(r, c) @ blk(b, c).T
        S = blk(r, b) - acc
        if b == r - 1:
            for c in range(b):
                accD += blk(r, c) @ blk(r, c).T
        if wait(diagf, b, 1) & FAILED:
            bail(); return
        blk(r, b)[:] = S @ W[b].T
        rowcnt[r] = b + 1
        if b == r - 1:
            accD += blk(r, b) @ blk(r, b).T
    D = np.tril(blk(r, r) - accD)
    D = D + np.tril(D, -1).T
    fail = 0
    L = np.zeros((FB, FB))
    for j in range(FB):                                       # right-looking, as chol64_smem
        d = D[j, j]
        if fail == 0 and not d > 0:
            fail = j + 1
        with np.errstate(invalid="ignore", divide="ignore"):
            L[j:, j] = D[j:, j] / np.sqrt(d); L[j, j] = np.sqrt(d)
            D[j + 1:, j + 1:] -= np.outer(L[j + 1:, j], L[j + 1:, j])
    if fail:
        if info[0] == 0:
            info[0] = r * FB + fail
        bail(); return
    if r + 1 < nblk:
        W[r] = np.linalg.inv(L)
        diagf[r] = 1
    blk(r, r)[:] = L + np.triu(blk(r, r), 1)


def run(nblk, bad, seed):
    rng = np.random.default_rng(seed)
    n = nblk * FB
    G = rng.random((n, n)); S = G @ G.T + n * np.eye(n)
    if bad is not None:
        S[bad, bad] = -1.0
    A = S.copy()
    W = [None] * nblk
    rowcnt, diagf, info = [0] * nblk, [0] * nblk, [0]
    errs, threads = [], []
    deadline = time.time() + 30
    order = list(range(nblk)); random.Random(seed).shuffle(order)      # CTAs do not start in index order
    for r in order:
        def body(r=r):
            try:
                cta(r, nblk, A, W, rowcnt, diagf, info, deadline, random.Random(seed * 100 + r))
            except Exception as ex:  # noqa: BLE001
                errs.append(ex)
        t = threading.Thread(target=body); threads.append(t); t.start()
    for t in threads:
        t.join()
    assert not errs, errs
    if bad is None:
        assert info[0] == 0
        ref = np.linalg.cholesky(S)
        assert np.abs(np.tril(A) - ref).max() < 1e-12 * np.abs(ref).max()
        assert np.array_equal(np.triu(A, 1), np.triu(S, 1))
    else:
        assert info[0] == bad + 1, (info[0], bad)


def main():
    for nblk in (1, 2, 5, 8, 16):
        for seed in range(3):
            run(nblk, None, seed)
    for nblk, bad in ((8, 0), (8, 37), (8, 63), (5, 20), (2, 9)):
        for seed in range(3):
            run(nblk, bad, seed)
    print("fused tile Cholesky flag protocol: no deadlock, L == LAPACK, first failing minor reported, FAILED propagates")


if __name__ == "__main__":
    main()
