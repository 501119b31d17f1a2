"""CPU emulation of the INDEX ARITHMETIC of slate_b200/csrc/potrf_tile_fused.cu (no GPU in the build container):
the loaders (load_block / load_block_t), the DMMA.8x8x4 and FP32 product micro-kernels (Prod<double>, Prod<float>:
per-thread fragment addresses and accumulator -> (row, col) maps), and the three block algorithms built on them
(tile Cholesky, Right/Lower/Trans panel solve, Left/Lower/NoTrans row solve) are transcribed thread by thread into
numpy and checked against LAPACK-style references, ragged sizes included.  This checks the transcription of the
algorithm and its indexing, not the CUDA code itself (barriers, flags, memory ordering are not modelled).
usage: python scratch/emulate_fused.py"""
import numpy as np

FB, FLD, FT = 64, 68, 128


class ProdD:
    def __init__(self, tid):
        warp, lane = tid >> 5, tid & 31
        self.lr, self.lc = lane >> 2, lane & 3
        self.wm, self.wn = (warp & 1) * 32, (warp >> 1) * 32

    def row(self, e): return self.wm + (e >> 3) * 8 + self.lr
    def col(self, e): return self.wn + ((e >> 1) & 3) * 8 + 2 * self.lc + (e & 1)


class ProdS:
    def __init__(self, tid):
        self.tx, self.ty = tid & 15, tid >> 4

    def row(self, e): return self.tx * 4 + (e >> 3)
    def col(self, e): return self.ty * 8 + (e & 7)


def mma_d(acc, Xs, Ys):
    """acc[tid][e]; warp-collective DMMA: D(8x8) += A(8x4) B(4x8); lane l supplies A[l/4][l%4], B[l%4][l/4] and
    receives C[l/4][2*(l%4) + {0,1}]"""
    for warp in range(FT // 32):
        wm, wn = (warp & 1) * 32, (warp >> 1) * 32
        for k4 in range(FB // 4):
            for i in range(4):
                for j in range(4):
                    Am = np.zeros((8, 4)); Bm = np.zeros((4, 8))
                    for lane in range(32):
                        lr, lc = lane >> 2, lane & 3
                        a = Xs[(lc + k4 * 4) * FLD + wm + lr + i * 8]       # cA[k4*4*FLD + i*8], cA = Xs + lc*FLD + wm + lr
                        b = Ys[(lc + k4 * 4) * FLD + wn + lr + j * 8]
                        Am[lane >> 2][lane & 3] = a
                        Bm[lane & 3][lane >> 2] = b
                    D = Am @ Bm
                    for lane in range(32):
                        tid = warp * 32 + lane
                        acc[tid][(i * 4 + j) * 2] += D[lane >> 2][2 * (lane & 3)]
                        acc[tid][(i * 4 + j) * 2 + 1] += D[lane >> 2][2 * (lane & 3) + 1]


def mma_s(acc, Xs, Ys):
    for tid in range(FT):
        tx, ty = tid & 15, tid >> 4
        for k in range(FB):
            a = Xs[k * FLD + tx * 4: k * FLD + tx * 4 + 4]
            b = Ys[k * FLD + ty * 8: k * FLD + ty * 8 + 8]
            for i in range(4):
                for j in range(8):
                    acc[tid][i * 8 + j] += a[i] * b[j]


def load_block(dst, mem, src, lds, rv):
    for tid in range(FT):
        i, k0 = tid & (FB - 1), tid >> 6
        for half in range(2):
            for t in range(FB // 4):
                k = k0 + 2 * (half * (FB // 4) + t)
                dst[k * FLD + i] = mem[src + i + k * lds] if i < rv else 0.0


def load_block_t(dst, mem, src, lds, cv):
    for tid in range(FT):
        k, j0 = tid & (FB - 1), tid >> 6
        for half in range(2):
            for t in range(FB // 4):
                j = j0 + 2 * (half * (FB // 4) + t)
                dst[k * FLD + j] = mem[src + k + j * lds] if j < cv else 0.0


def chol64_smem(As, Ls, rd, NT, LD, rsq):
    """diag64.cuh: chol64_smem, thread by thread; phases separated by the kernel's barriers"""
    TC = NT // 16
    fail = 0
    for j in range(64):
        d = As[j * LD + j]
        if fail == 0 and not d > 0:
            fail = j + 1
        if rsq:
            rinv = 1.0 / np.sqrt(d); diag = d * rinv
        else:
            diag = np.sqrt(d); rinv = 1.0 / diag
        for tid in range(NT):                                   # scale phase
            for r in range(j + tid, 64, NT):
                Ls[j * LD + r] = diag if r == j else As[j * LD + r] * rinv
            if tid == 0:
                rd[j] = rinv
        for tid in range(NT):                                   # update phase
            tr, tc = tid & 15, tid >> 4
            for c in range(j + 1 + tc, 64, TC):
                lc = Ls[j * LD + c]
                for r in range(c + tr, 64, 16):
                    As[c * LD + r] -= Ls[j * LD + r] * lc
    return fail


def inv64_smem(Ls, rd, Xs, NT, LD):
    TC = NT // 16
    for tid in range(NT):
        for e in range(tid, 64 * 64, NT):
            r, c = e & 63, e >> 6
            Xs[c * LD + r] = 1.0 if r == c else 0.0
    for i in range(64):
        for tid in range(NT):
            if tid <= i:
                Xs[tid * LD + i] *= rd[i]
        for tid in range(NT):
            tr, tc = tid & 15, tid >> 4
            for j in range(tc, i + 1, TC):
                xij = Xs[j * LD + i]
                for r in range(i + 1 + tr, 64, 16):
                    Xs[j * LD + r] -= Ls[i * LD + r] * xij


def trsm_lln_small(B, na, n, ldb, alpha, T, ldt, unit, NT=256):
    """potrf_tile_fused.cu: trsm_lln_small_kernel (one CTA per 64 columns), phases separated by its barriers"""
    for c0 in range(0, n, 64):
        cv = min(64, n - c0)
        Ls = np.zeros(64 * 65); Bs = np.zeros(64 * 65)
        for e in range(na * na):
            i, k = e % na, e // na
            if i > k:
                Ls[k * 65 + i] = T[i + k * ldt]
            elif i == k:
                Ls[k * 65 + i] = 1.0 if unit else 1.0 / T[i + k * ldt]
        for e in range(na * cv):
            i, c = e % na, e // na
            Bs[c * 65 + i] = alpha * B[i + (c0 + c) * ldb]
        for k in range(na):
            if not unit:
                for tid in range(NT):
                    if tid < cv:
                        Bs[tid * 65 + k] *= Ls[k * 65 + k]
            for tid in range(NT):
                ti, tcol = tid & 15, tid >> 4
                for c in range(tcol, cv, NT // 16):
                    xk = Bs[c * 65 + k]
                    for i in range(k + 1 + ti, na, 16):
                        Bs[c * 65 + i] -= Ls[k * 65 + i] * xk
        for e in range(na * cv):
            i, c = e % na, e // na
            B[i + (c0 + c) * ldb] = Bs[c * 65 + i]


def make(prec):
    P = ProdD if prec == "d" else ProdS
    mma = mma_d if prec == "d" else mma_s
    return [P(t) for t in range(FT)], mma


def potrf_tile(A, n, lda, prec):
    """A: flat column-major array (modified in place).  CTAs run in index order (the flags only order them)."""
    pr, mma = make(prec)
    nblk = -(-n // FB)
    W = np.zeros(nblk * FB * FB)
    Xs = np.zeros(FB * FLD); Ys = np.zeros(FB * FLD); Cs = np.zeros(FB * FLD)
    for r in range(nblk):
        rv = min(FB, n - r * FB)
        Arow = r * FB
        accD = np.zeros((FT, 32))
        for b in range(r):
            acc = np.zeros((FT, 32))
            for c in range(b):
                load_block(Xs, A, Arow + c * FB * lda, lda, rv)
                load_block(Ys, A, b * FB + c * FB * lda, lda, FB)
                mma(acc, Xs, Ys)
            Ab = Arow + b * FB * lda
            for tid in range(FT):
                for e in range(32):
                    row, col = pr[tid].row(e), pr[tid].col(e)
                    o = A[Ab + row + col * lda] if row < rv else 0.0
                    Cs[col * FLD + row] = o - acc[tid][e]
            if b == r - 1:
                for c in range(b):
                    load_block(Xs, A, Arow + c * FB * lda, lda, rv)
                    mma(accD, Xs, Xs)
            load_block(Ys, W, b * FB * FB, FB, FB)
            acc = np.zeros((FT, 32))
            mma(acc, Cs, Ys)
            for tid in range(FT):
                for e in range(32):
                    row, col = pr[tid].row(e), pr[tid].col(e)
                    if row < rv:
                        A[Ab + row + col * lda] = acc[tid][e]
                    if b == r - 1:
                        Xs[col * FLD + row] = acc[tid][e]
            if b == r - 1:
                mma(accD, Xs, Xs)
        Ad = Arow + r * FB * lda
        for tid in range(FT):
            for e in range(32):
                row, col = pr[tid].row(e), pr[tid].col(e)
                if row < rv and col < rv:
                    v = A[Ad + row + col * lda] - accD[tid][e] if col <= row else 0.0
                else:
                    v = 1.0 if row == col else 0.0
                Cs[col * FLD + row] = v
        rd = np.zeros(FB)
        fail = chol64_smem(Cs, Xs, rd, FT, FLD, False)            # D in Cs -> L in Xs
        assert fail == 0
        for e in range(FB * FB):
            row, col = e & (FB - 1), e >> 6
            if col <= row and row < rv:
                A[Ad + row + col * lda] = Xs[col * FLD + row]
        inv64_smem(Xs, rd, Ys, FT, FLD)
        for e in range(FB * FB):
            W[r * FB * FB + e] = Ys[(e >> 6) * FLD + (e & (FB - 1))]


def trsm_rlt(B, m, na, ldb, alpha, T, ldt, prec):
    pr, mma = make(prec)
    nblk = -(-na // FB)
    W = np.zeros(nblk * FB * FB)
    for j in range(nblk):
        jv = min(FB, na - j * FB)
        Lj = np.eye(FB)
        for a in range(jv):
            for b in range(a + 1):
                Lj[a, b] = T[j * FB + a + (j * FB + b) * ldt]
        Wj = np.linalg.inv(Lj)
        for a in range(FB):
            for b in range(FB):
                W[j * FB * FB + a + b * FB] = Wj[a, b]
    Xs = np.zeros(FB * FLD); Ys = np.zeros(FB * FLD); Cs = np.zeros(FB * FLD)
    for rb in range(-(-m // FB)):
        r0 = rb * FB
        rv = min(FB, m - r0)
        Brow = r0
        for j in range(nblk):
            jv = min(FB, na - j * FB)
            acc = np.zeros((FT, 32))
            for c in range(j):
                load_block(Xs, B, Brow + c * FB * ldb, ldb, rv)
                load_block(Ys, T, j * FB + c * FB * ldt, ldt, jv)
                mma(acc, Xs, Ys)
            Bj = Brow + j * FB * ldb
            for tid in range(FT):
                for e in range(32):
                    row, col = pr[tid].row(e), pr[tid].col(e)
                    o = B[Bj + row + col * ldb] if (row < rv and col < jv) else 0.0
                    Cs[col * FLD + row] = alpha * o - acc[tid][e]
            load_block(Ys, W, j * FB * FB, FB, FB)
            acc = np.zeros((FT, 32))
            mma(acc, Cs, Ys)
            for tid in range(FT):
                for e in range(32):
                    row, col = pr[tid].row(e), pr[tid].col(e)
                    if row < rv and col < jv:
                        B[Bj + row + col * ldb] = acc[tid][e]


def trsm_lln(B, na, n, ldb, alpha, T, ldt, unit, prec):
    pr, mma = make(prec)
    nblk = -(-na // FB)
    W = np.zeros(nblk * FB * FB)
    for j in range(nblk):
        jv = min(FB, na - j * FB)
        Lj = np.eye(FB)
        for a in range(jv):
            for b in range(a + 1):
                if a != b or not unit:
                    Lj[a, b] = T[j * FB + a + (j * FB + b) * ldt]
        Wj = np.linalg.inv(Lj)
        for a in range(FB):
            for b in range(FB):
                W[j * FB * FB + a + b * FB] = Wj[a, b]
    Xs = np.zeros(FB * FLD); Ys = np.zeros(FB * FLD); Cs = np.zeros(FB * FLD)
    for cbk in range(-(-n // FB)):
        c0 = cbk * FB
        cv = min(FB, n - c0)
        Bcol = c0 * ldb
        for j in range(nblk):
            jv = min(FB, na - j * FB)
            acc = np.zeros((FT, 32))
            for c in range(j):
                load_block(Xs, T, j * FB + c * FB * ldt, ldt, jv)
                load_block_t(Ys, B, Bcol + c * FB, ldb, cv)
                mma(acc, Xs, Ys)
            Bj = Bcol + j * FB
            for tid in range(FT):
                for e in range(32):
                    row, col = pr[tid].row(e), pr[tid].col(e)
                    o = B[Bj + row + col * ldb] if (row < jv and col < cv) else 0.0
                    Cs[row * FLD + col] = alpha * o - acc[tid][e]
            load_block(Xs, W, j * FB * FB, FB, FB)
            acc = np.zeros((FT, 32))
            mma(acc, Xs, Cs)
            for tid in range(FT):
                for e in range(32):
                    row, col = pr[tid].row(e), pr[tid].col(e)
                    if row < jv and col < cv:
                        B[Bj + row + col * ldb] = acc[tid][e]


def main():
    rng = np.random.default_rng(1)
    # stand-alone multi-warp diagonal kernels (256 threads, LD = 65): L and the full inverse, both forms of the pivot
    for rsq in (False, True):
        G = rng.random((64, 64)); S = G @ G.T + 64 * np.eye(64)
        LD = 65
        As = np.zeros(64 * LD); Ls = np.zeros(64 * LD); rd = np.zeros(64); Xs = np.full(64 * LD, 3.0)
        for c in range(64):
            for r in range(64):
                As[c * LD + r] = S[r, c] if c <= r else 0.0
        assert chol64_smem(As, Ls, rd, 256, LD, rsq) == 0
        L = np.array([[Ls[c * LD + r] if c <= r else 0.0 for c in range(64)] for r in range(64)])
        ref = np.linalg.cholesky(S)
        assert np.abs(L - ref).max() < 1e-13 * np.abs(ref).max()
        inv64_smem(Ls, rd, Xs, 256, LD)
        X = np.array([[Xs[c * LD + r] for c in range(64)] for r in range(64)])
        assert np.abs(X - np.linalg.inv(ref)).max() < 1e-12 * np.abs(X).max() and (np.triu(X, 1) == 0).all()
    S = np.eye(64); S[40, 40] = -1.0
    As = np.zeros(64 * 65)
    for c in range(64):
        for r in range(c, 64):
            As[c * 65 + r] = S[r, c]
    with np.errstate(invalid="ignore"):
        assert chol64_smem(As, np.zeros(64 * 65), np.zeros(64), 256, 65, False) == 41
    print("chol64_smem / inv64_smem (multi-warp diagonal block): OK")
    for na, n, unit in ((32, 32, True), (64, 100, False), (20, 7, True), (33, 130, False)):
        T = rng.random((na, na)) / na + np.eye(na) * (1 + rng.random(na))
        Bm = rng.random((na, n))
        flat = Bm.flatten(order="F")
        trsm_lln_small(flat, na, n, na, 0.7, T.flatten(order="F"), na, unit)
        Lm = np.tril(T, -1) + (np.eye(na) if unit else np.diag(np.diag(T)))
        ref = 0.7 * np.linalg.solve(Lm, Bm)
        assert np.abs(flat.reshape((na, n), order="F") - ref).max() < 1e-13 * np.abs(ref).max(), (na, n, unit)
    print("trsm_lln_small_kernel (direct substitution, small triangle): OK")
    for prec in ("d", "s"):
        # every (row, col) of the 64 x 64 block is owned by exactly one accumulator
        pr, _ = make(prec)
        seen = np.zeros((FB, FB), int)
        for t in range(FT):
            for e in range(32):
                seen[pr[t].row(e), pr[t].col(e)] += 1
        assert (seen == 1).all(), prec
        for n, lda in ((130, 136), (192, 192)):
            G = rng.random((n, n)); S = G @ G.T + n * np.eye(n)
            buf = np.full((lda, n), 7.25); buf[:n] = S
            flat = buf.flatten(order="F")
            potrf_tile(flat, n, lda, prec)
            out = flat.reshape((lda, n), order="F")
            ref = np.linalg.cholesky(S)
            assert np.abs(np.tril(out[:n]) - ref).max() < 1e-12 * np.abs(ref).max(), (prec, n)
            assert np.array_equal(np.triu(out[:n], 1), np.triu(S, 1)) and (out[n:] == 7.25).all()
        for m, na in ((70, 130), (64, 192)):
            T = rng.random((na, na)) / na + np.eye(na) * (1 + rng.random(na))
            Bm = rng.random((m, na))
            flat = Bm.flatten(order="F")
            trsm_rlt(flat, m, na, m, 0.7, T.flatten(order="F"), na, prec)
            ref = 0.7 * Bm @ np.linalg.inv(np.tril(T)).T
            assert np.abs(flat.reshape((m, na), order="F") - ref).max() < 1e-12 * np.abs(ref).max(), (prec, m, na)
        for na, n, unit in ((130, 70, True), (192, 64, False)):
            T = rng.random((na, na)) / na + np.eye(na) * (1 + rng.random(na))
            Bm = rng.random((na, n))
            flat = Bm.flatten(order="F")
            trsm_lln(flat, na, n, na, 0.7, T.flatten(order="F"), na, unit, prec)
            Lm = np.tril(T, -1) + (np.eye(na) if unit else np.diag(np.diag(T)))
            ref = 0.7 * np.linalg.solve(Lm, Bm)
            assert np.abs(flat.reshape((na, n), order="F") - ref).max() < 1e-12 * np.abs(ref).max(), (prec, na, n)
        print(f"Prod<{prec}>: tile Cholesky, panel solve and row solve index arithmetic OK")


if __name__ == "__main__":
    main()
