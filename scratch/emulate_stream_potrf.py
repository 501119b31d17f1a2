"""CPU model of the schedule of the streaming-input Cholesky (potrf_driver with host_in, slate_b200/csrc/runtime.cu):
the driver's loop is transcribed as a list of operations per stream with its event waits, executed by a tiny
two-stream + copy-stream simulator on numpy tiles in EVERY order the events allow (random interleavings), and compared
with the plain right-looking tile Cholesky: bitwise equal factors (every tile sees the same updates in the same order),
no operation touches a chunk before it has "arrived".
usage: python scratch/emulate_stream_potrf.py"""
import random

import numpy as np


def chunks(nt, cw):
    cb = [0]
    c = min(nt, max(1, cw // 2))
    while c < nt:
        cb.append(c); c += cw
    cb.append(nt)
    return cb


def build(nt, cw):
    """returns streams: dict name -> list of ops; op = (kind, payload, waits, records)"""
    cb = chunks(nt, cw)
    nchunk = len(cb) - 1
    chunk_of = [0] * nt
    for c in range(nchunk):
        for j in range(cb[c], cb[c + 1]):
            chunk_of[j] = c
    P, T, C = [], [], []
    for c in range(nchunk):
        C.append(("h2d", c, [], [("H", c)]))
    la = {k: [(i, k + 1) for i in range(k + 1, nt)] if k + 1 < nt and chunk_of[k + 1] == chunk_of[k] else [] for k in range(nt)}
    tr = {k: [(i, j) for j in range(k + 2, nt) if chunk_of[j] == chunk_of[k] for i in range(j, nt)] for k in range(nt)}
    cu = {(k, c): [(i, j) for j in range(k + 1, nt) if chunk_of[j] == c and c > chunk_of[k] for i in range(j, nt)]
          for k in range(nt) for c in range(nchunk)}
    for k in range(nt):
        c = chunk_of[k]
        if k == cb[c]:
            T.append(("wait", None, [("H", c)], []))
            P.append(("wait", None, [("H", c)], []))
            if c > 0:
                T.append(("wait", None, [("P", k - 1)], []))
                for kk in range(k):
                    if cu[(kk, c)]:
                        T.append(("update", (kk, cu[(kk, c)]), [], []))
                T.append(("rec", None, [], [("C", c)]))
                P.append(("wait", None, [("C", c)], []))
        if k >= 1 and la[k - 1]:
            w = [("T", k - 2)] if k >= 2 else []
            P.append(("update", (k - 1, la[k - 1]), w, []))
        P.append(("potrf", k, [], []))
        if k + 1 < nt:
            P.append(("trsm", k, [], []))
        P.append(("rec", None, [], [("P", k)]))
        T.append(("wait", None, [("P", k)], []))
        if tr[k]:
            T.append(("update", (k, tr[k]), [], []))
        T.append(("rec", None, [], [("T", k)]))
    return {"P": P, "T": T, "C": C}, cb, chunk_of


def run(nt, nb, cw, seed):
    rng = np.random.default_rng(seed)
    n = nt * nb
    G = rng.random((n, n)); S = G @ G.T + n * np.eye(n)
    tile = lambda M, i, j: M[i * nb:(i + 1) * nb, j * nb:(j + 1) * nb]
    # reference: plain right-looking
    R = S.copy()
    for k in range(nt):
        tile(R, k, k)[:] = np.linalg.cholesky(tile(R, k, k))
        for i in range(k + 1, nt):
            tile(R, i, k)[:] = np.linalg.solve(tile(R, k, k), tile(R, i, k).T).T
        for j in range(k + 1, nt):
            for i in range(j, nt):
                tile(R, i, j)[:] -= tile(R, i, k) @ tile(R, j, k).T
    streams, cb, chunk_of = build(nt, cw)
    A = np.full((n, n), np.nan)                                # device pool: nothing has arrived yet
    done = set()
    pos = {s: 0 for s in streams}
    rnd = random.Random(seed)
    while any(pos[s] < len(streams[s]) for s in streams):
        ready = [s for s in streams if pos[s] < len(streams[s]) and all(w in done for w in streams[s][pos[s]][2])]
        assert ready, "deadlock in the modelled schedule"
        s = rnd.choice(ready)
        kind, pay, _, recs = streams[s][pos[s]]
        pos[s] += 1
        if kind == "h2d":
            for j in range(cb[pay], cb[pay + 1]):
                A[j * nb:, j * nb:(j + 1) * nb] = S[j * nb:, j * nb:(j + 1) * nb]
        elif kind == "potrf":
            t = tile(A, pay, pay); assert not np.isnan(np.tril(t)).any()
            t[:] = np.linalg.cholesky(np.tril(t) + np.tril(t, -1).T)
        elif kind == "trsm":
            for i in range(pay + 1, nt):
                t = tile(A, i, pay); assert not np.isnan(t).any()
                t[:] = np.linalg.solve(tile(A, pay, pay), t.T).T
        elif kind == "update":
            k, lst = pay
            for (i, j) in lst:
                t = tile(A, i, j); assert not np.isnan(np.tril(t) if i == j else t).any(), (k, i, j)
                t[:] -= tile(A, i, k) @ tile(A, j, k).T
        for r in recs:
            done.add(r)
    L, Lr = np.tril(A), np.tril(R)
    assert np.array_equal(L, Lr), np.abs(L - Lr).max()


def main():
    for nt, cw in ((9, 4), (6, 1), (12, 8), (5, 2), (3, 8), (1, 8)):
        for seed in range(3):
            run(nt, 8, cw, seed)
        print(f"streaming potrf schedule: nt={nt} chunk={cw}: bitwise the right-looking factor in 3 random interleavings")


if __name__ == "__main__":
    main()
