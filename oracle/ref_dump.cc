// oracle/ref_dump.cc -- TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// A small driver over the UNMODIFIED reference library (oracle/_ref/libslate_ref.so,
// built by oracle/build_ref.sh).  For one (routine, type, n, nb, seeds) it
//   1. generates the inputs with the reference's own matgen (Philox-2x64 keyed on
//      global (i, j, seed): matgen/random.cc:54-159),
//   2. runs the reference's Target::HostTask path through the same simplified-API
//      calls the reference tester makes (test/test_gemm.cc:177, test_posv.cc:211,
//      test_gesv.cc:233, test_herk.cc, test_trsm.cc),
//   3. writes inputs / outputs / pivots as raw little-endian column-major arrays and
//      prints one JSON line with the wall time of the routine.
// The numpy restatement in oracle/slate_oracle.py and the CUDA path are both compared
// against these files (tests/golden/ holds the small committed ones; the script that
// produced them is tests/golden/make_golden.py).
//
// usage: ref_dump ROUTINE TYPE n nb seedA seedB seedC OUTPREFIX [key=value ...]
//   ROUTINE  gen | gemm | herk | potrf | getrf | trsm | gesv_mixed | posv_mixed | posv | gesv | hemm | norms
//   TYPE     s | d | c | z
//   keys     p=1 q=1 (process grid; > 1 rank: oracle/_ref/ref_dump_mp under oracle/mprun.py)  kind=rand|rand_dominant  la=1  ib=16  threads=N  dump=0|1  nrhs=10  pt=panel threads  method=pplu|calu (getrf)
//            m= k= (gemm/herk rectangular)  uplo=l|u
#include "slate/slate.hh"
#include "slate/generate_matrix.hh"
#include "lapack/flops.hh"

#include <omp.h>
#include <chrono>
#include <cmath>
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <string>
#include <vector>

namespace {

using Clock = std::chrono::steady_clock;

// process grid (keys p=, q=; the world size has to be p * q).  More than one rank needs the multi-process MPI replacement
// (oracle/mpi_mp, started through oracle/mprun.py); every rank then writes the tiles IT owns (other tiles stay zero) to
// PREFIX.r<rank>.<name>.bin and the caller assembles them with the tile map  rank(i, j) = i % p + (j % q) * p.
int g_p = 1, g_q = 1, g_rank = 0, g_world = 1;

struct Args {
    std::string routine, type, prefix;
    int64_t n = 0, nb = 0, seedA = 42, seedB = 43, seedC = 44;
    std::map<std::string, std::string> kv;
    std::string get(const std::string& k, const std::string& d) const {
        auto it = kv.find(k);
        return it == kv.end() ? d : it->second;
    }
    int64_t geti(const std::string& k, int64_t d) const {
        auto it = kv.find(k);
        return it == kv.end() ? d : std::atoll(it->second.c_str());
    }
};

template <typename T>
void write_raw(const std::string& path, const T* data, size_t count)
{
    FILE* f = std::fopen(path.c_str(), "wb");
    if (! f) { std::perror(path.c_str()); std::exit(2); }
    std::fwrite(data, sizeof(T), count, f);
    std::fclose(f);
}

// Gather a general tiled matrix into one column-major m-by-n array.
template <typename T>
std::vector<T> to_dense(slate::Matrix<T>& A)
{
    int64_t m = A.m(), n = A.n();
    std::vector<T> out(size_t(m) * n);
    int64_t j0 = 0;
    for (int64_t j = 0; j < A.nt(); ++j) {
        int64_t i0 = 0;
        for (int64_t i = 0; i < A.mt(); ++i) {
            if (A.tileIsLocal(i, j)) {
                A.tileGetForReading(i, j, slate::LayoutConvert::ColMajor);
                auto t = A(i, j);
                for (int64_t jj = 0; jj < t.nb(); ++jj)
                    for (int64_t ii = 0; ii < t.mb(); ++ii)
                        out[size_t(i0 + ii) + size_t(j0 + jj) * m] = t(ii, jj);
            }
            i0 += A.tileMb(i);
        }
        j0 += A.tileNb(j);
    }
    return out;
}

// Gather the stored triangle of a Hermitian/triangular matrix (other triangle = 0).
template <typename MatrixT, typename T>
std::vector<T> tz_to_dense(MatrixT& A, bool lower)
{
    int64_t n = A.n();
    std::vector<T> out(size_t(n) * n, T(0));
    int64_t j0 = 0;
    for (int64_t j = 0; j < A.nt(); ++j) {
        int64_t i0 = 0;
        for (int64_t i = 0; i < A.mt(); ++i) {
            bool stored = lower ? (i >= j) : (i <= j);
            if (stored && A.tileIsLocal(i, j)) {
                A.tileGetForReading(i, j, slate::LayoutConvert::ColMajor);
                auto t = A(i, j);
                for (int64_t jj = 0; jj < t.nb(); ++jj)
                    for (int64_t ii = 0; ii < t.mb(); ++ii)
                        out[size_t(i0 + ii) + size_t(j0 + jj) * n] = t(ii, jj);
            }
            i0 += A.tileMb(i);
        }
        j0 += A.tileNb(j);
    }
    return out;
}

template <typename T>
slate::Matrix<T> make_matrix(int64_t m, int64_t n, int64_t nb, int64_t seed, const std::string& kind)
{
    slate::Matrix<T> A(m, n, nb, nb, slate::GridOrder::Col, g_p, g_q, MPI_COMM_WORLD);
    A.insertLocalTiles();
    slate::MatgenParams p;
    p.verbose = 0; p.kind = kind; p.cond_request = NAN; p.cond_actual = NAN; p.condD = NAN; p.seed = seed;
    slate::generate_matrix(p, A);
    return A;
}

template <typename T>
int run(const Args& a)
{
    using real_t = blas::real_type<T>;
    const int64_t n = a.n, nb = a.nb;
    const bool dump = a.geti("dump", 1) != 0;
    const int64_t la = a.geti("la", 1);
    const int64_t ib = a.geti("ib", 16);
    const int64_t nrhs = a.geti("nrhs", 10);
    const int64_t pt = a.geti("pt", std::max(omp_get_max_threads() / 2, 1));
    slate::Options opts = {
        {slate::Option::Lookahead, la},
        {slate::Option::Target, slate::Target::HostTask},
        {slate::Option::InnerBlocking, ib},
        {slate::Option::MaxPanelThreads, pt},
    };
    double seconds = 0, gflop = 0;
    int64_t info = 0;
    int iters = 0;
    auto tic = [] { return Clock::now(); };
    auto toc = [](Clock::time_point t0) {
        return std::chrono::duration<double>(Clock::now() - t0).count(); };

    // tester defaults: alpha = pi + sqrt(2) i, beta = e + sqrt(3) i  (test/test.cc:447-448)
    T alpha = blas::make_scalar<T>(real_t(3.141592653589793), real_t(1.414213562373095));
    T beta  = blas::make_scalar<T>(real_t(2.718281828459045), real_t(1.732050807568877));

    if (a.routine == "gen") {
        std::string kind = a.get("kind", "rand");
        auto A = make_matrix<T>(a.geti("m", n), n, nb, a.seedA, kind);
        auto d = to_dense(A);
        write_raw(a.prefix + ".A.bin", d.data(), d.size());
    }
    else if (a.routine == "gemm") {
        // C = alpha op(A) op(B) + beta C; opa / opb = n|t|c: A is stored k x m / B n x k and handed over as a (conjugate-)
        // transposed view (test/test_gemm.cc:96-135)
        int64_t m = a.geti("m", n), k = a.geti("k", n);
        const std::string opa = a.get("opa", "n"), opb = a.get("opb", "n");
        auto A = opa == "n" ? make_matrix<T>(m, k, nb, a.seedA, "rand") : make_matrix<T>(k, m, nb, a.seedA, "rand");
        auto B = opb == "n" ? make_matrix<T>(k, n, nb, a.seedB, "rand") : make_matrix<T>(n, k, nb, a.seedB, "rand");
        auto C = make_matrix<T>(m, n, nb, a.seedC, "rand");
        if (dump) {
            auto d = to_dense(A); write_raw(a.prefix + ".A.bin", d.data(), d.size());
            d = to_dense(B);      write_raw(a.prefix + ".B.bin", d.data(), d.size());
            d = to_dense(C);      write_raw(a.prefix + ".C.bin", d.data(), d.size());
        }
        auto opA = A, opB = B;
        if (opa == "t") opA = slate::transpose(A); else if (opa == "c") opA = slate::conj_transpose(A);
        if (opb == "t") opB = slate::transpose(B); else if (opb == "c") opB = slate::conj_transpose(B);
        auto t0 = tic();
        slate::multiply(alpha, opA, opB, beta, C, opts);
        seconds = toc(t0);
        gflop = blas::Gflop<T>::gemm(m, n, k);
        if (dump) { auto d = to_dense(C); write_raw(a.prefix + ".out.bin", d.data(), d.size()); }
    }
    else if (a.routine == "herk") {
        // trans=c: A is stored k x n and handed over as its conjugate-transposed view, C = alpha A^H A + beta C (test/test_herk.cc)
        int64_t k = a.geti("k", n);
        bool lower = a.get("uplo", "l") == "l";
        const bool tr = a.get("trans", "n") != "n";
        auto A = tr ? make_matrix<T>(k, n, nb, a.seedA, "rand") : make_matrix<T>(n, k, nb, a.seedA, "rand");
        slate::HermitianMatrix<T> C(lower ? slate::Uplo::Lower : slate::Uplo::Upper, n, nb,
                                    slate::GridOrder::Col, g_p, g_q, MPI_COMM_WORLD);
        C.insertLocalTiles();
        slate::MatgenParams p; p.verbose = 0; p.kind = "rand"; p.seed = a.seedC;
        p.cond_request = p.cond_actual = p.condD = NAN;
        slate::generate_matrix(p, C);
        if (dump) {
            auto d = to_dense(A); write_raw(a.prefix + ".A.bin", d.data(), d.size());
            d = tz_to_dense<slate::HermitianMatrix<T>, T>(C, lower);
            write_raw(a.prefix + ".C.bin", d.data(), d.size());
        }
        real_t ra = std::real(alpha), rb = std::real(beta);
        auto t0 = tic();
        auto opA = A;
        if (tr) opA = slate::conj_transpose(A);
        slate::rank_k_update(ra, opA, rb, C, opts);
        seconds = toc(t0);
        gflop = blas::Gflop<T>::herk(n, k);
        if (dump) {
            auto d = tz_to_dense<slate::HermitianMatrix<T>, T>(C, lower);
            write_raw(a.prefix + ".out.bin", d.data(), d.size());
        }
    }
    else if (a.routine == "potrf") {
        bool lower = a.get("uplo", "l") == "l";
        slate::HermitianMatrix<T> A(lower ? slate::Uplo::Lower : slate::Uplo::Upper, n, nb,
                                    slate::GridOrder::Col, g_p, g_q, MPI_COMM_WORLD);
        A.insertLocalTiles();
        slate::MatgenParams p; p.verbose = 0; p.kind = a.get("kind", "rand_dominant"); p.seed = a.seedA;
        p.cond_request = p.cond_actual = p.condD = NAN;
        slate::generate_matrix(p, A);
        if (dump) {
            auto d = tz_to_dense<slate::HermitianMatrix<T>, T>(A, lower);
            write_raw(a.prefix + ".A.bin", d.data(), d.size());
        }
        auto t0 = tic();
        info = slate::chol_factor(A, opts);
        seconds = toc(t0);
        gflop = lapack::Gflop<T>::potrf(n);
        if (dump) {
            auto d = tz_to_dense<slate::HermitianMatrix<T>, T>(A, lower);
            write_raw(a.prefix + ".out.bin", d.data(), d.size());
        }
    }
    else if (a.routine == "getrf") {
        int64_t m = a.geti("m", n);
        auto A = make_matrix<T>(m, n, nb, a.seedA, a.get("kind", "rand"));
        if (dump) { auto d = to_dense(A); write_raw(a.prefix + ".A.bin", d.data(), d.size()); }
        slate::Pivots pivots;
        // method=calu: tournament pivoting (slate::getrf_tntpiv, src/getrf_tntpiv.cc) through the same lu_factor call
        // (Option::MethodLU, src/getrf.cc:324-329); one rank here, so the tournament has a single participant
        slate::Options lu_opts = opts;
        if (a.get("method", "pplu") == "calu") lu_opts[slate::Option::MethodLU] = slate::MethodLU::CALU;
        auto t0 = tic();
        info = slate::lu_factor(A, pivots, lu_opts);
        seconds = toc(t0);
        gflop = lapack::Gflop<T>::getrf(m, n);
        if (dump) {
            auto d = to_dense(A); write_raw(a.prefix + ".out.bin", d.data(), d.size());
            // pivots: for each block column k, min(mb,nb) pairs (tileIndex, elementOffset)
            // relative to the panel sub-matrix A(k:mt-1, k)   (include/slate/types.hh:84-105)
            std::vector<int64_t> flat;
            for (auto& col : pivots)
                for (auto& pv : col) { flat.push_back(pv.tileIndex()); flat.push_back(pv.elementOffset()); }
            write_raw(a.prefix + ".piv.bin", flat.data(), flat.size());
        }
    }
    else if (a.routine == "getrf_nopiv") {
        // LU without pivoting (slate::getrf_nopiv, src/getrf_nopiv.cc); the tester uses a diagonally dominant matrix
        auto A = make_matrix<T>(n, n, nb, a.seedA, a.get("kind", "rand_dominant"));
        auto t0 = tic();
        info = slate::getrf_nopiv(A, opts);
        seconds = toc(t0);
        gflop = lapack::Gflop<T>::getrf(n, n);
        if (dump) { auto d = to_dense(A); write_raw(a.prefix + ".out.bin", d.data(), d.size()); }
    }
    else if (a.routine == "trsm") {
        // op(A) X = alpha B (side=l, default: A m x m) or X op(A) = alpha B (side=r: A n x n), B m x n; A = rand_dominant
        // lower triangle, op=n|t|c as a transposed view, diag=n|u  (test/test_trsm.cc; slate::trsm, src/trsm.cc)
        int64_t m = a.geti("m", n);
        const bool right = a.get("side", "l") == "r";
        const bool unit = a.get("diag", "n") == "u";
        const std::string op = a.get("op", "n");
        slate::TriangularMatrix<T> A(slate::Uplo::Lower, unit ? slate::Diag::Unit : slate::Diag::NonUnit, right ? n : m, nb,
                                     slate::GridOrder::Col, g_p, g_q, MPI_COMM_WORLD);
        A.insertLocalTiles();
        slate::MatgenParams p; p.verbose = 0; p.kind = "rand_dominant"; p.seed = a.seedA;
        p.cond_request = p.cond_actual = p.condD = NAN;
        slate::generate_matrix(p, A);
        auto B = make_matrix<T>(m, n, nb, a.seedB, "rand");
        if (dump) {
            auto d = tz_to_dense<slate::TriangularMatrix<T>, T>(A, true);
            write_raw(a.prefix + ".A.bin", d.data(), d.size());
            d = to_dense(B); write_raw(a.prefix + ".B.bin", d.data(), d.size());
        }
        auto opA = A;
        if (op == "t") opA = slate::transpose(A);
        else if (op == "c") opA = slate::conj_transpose(A);
        auto t0 = tic();
        if (right) slate::trsm(slate::Side::Right, alpha, opA, B, opts);
        else       slate::triangular_solve(alpha, opA, B, opts);
        seconds = toc(t0);
        gflop = blas::Gflop<T>::trsm(right ? slate::Side::Right : slate::Side::Left, m, n);
        if (dump) { auto d = to_dense(B); write_raw(a.prefix + ".out.bin", d.data(), d.size()); }
    }
    else if (a.routine == "gesv_mixed") {
        if constexpr (std::is_same<real_t, double>::value) {
            auto A = make_matrix<T>(n, n, nb, a.seedA, a.get("kind", "rand"));
            auto B = make_matrix<T>(n, nrhs, nb, a.seedB, "rand");
            slate::Matrix<T> X(n, nrhs, nb, nb, slate::GridOrder::Col, g_p, g_q, MPI_COMM_WORLD);
            X.insertLocalTiles();
            if (dump) {
                auto d = to_dense(A); write_raw(a.prefix + ".A.bin", d.data(), d.size());
                d = to_dense(B);      write_raw(a.prefix + ".B.bin", d.data(), d.size());
            }
            slate::Pivots pivots;
            auto t0 = tic();
            info = slate::gesv_mixed(A, pivots, B, X, iters, opts);
            seconds = toc(t0);
            gflop = lapack::Gflop<T>::gesv(n, nrhs);
            if (dump) { auto d = to_dense(X); write_raw(a.prefix + ".out.bin", d.data(), d.size()); }
        }
        else {
            std::fprintf(stderr, "gesv_mixed needs type d or z\n");
            return 2;
        }
    }
    else if (a.routine == "posv_mixed" || a.routine == "posv") {
        // Hermitian positive definite solve: posv = chol_factor + chol_solve_using_factor
        // (test/test_posv.cc:205-232); posv_mixed = src/posv_mixed.cc
        slate::HermitianMatrix<T> A(slate::Uplo::Lower, n, nb, slate::GridOrder::Col, g_p, g_q, MPI_COMM_WORLD);
        A.insertLocalTiles();
        slate::MatgenParams p; p.verbose = 0; p.kind = a.get("kind", "rand_dominant"); p.seed = a.seedA;
        p.cond_request = p.cond_actual = p.condD = NAN;
        slate::generate_matrix(p, A);
        auto B = make_matrix<T>(n, nrhs, nb, a.seedB, "rand");
        if (a.routine == "posv") {
            auto t0 = tic();
            info = slate::chol_factor(A, opts);
            if (info == 0) slate::chol_solve_using_factor(A, B, opts);
            seconds = toc(t0);
            gflop = lapack::Gflop<T>::posv(n, nrhs);
            if (dump) { auto d = to_dense(B); write_raw(a.prefix + ".out.bin", d.data(), d.size()); }
        }
        else if constexpr (std::is_same<real_t, double>::value) {
            slate::Matrix<T> X(n, nrhs, nb, nb, slate::GridOrder::Col, g_p, g_q, MPI_COMM_WORLD);
            X.insertLocalTiles();
            auto t0 = tic();
            info = slate::posv_mixed(A, B, X, iters, opts);
            seconds = toc(t0);
            gflop = lapack::Gflop<T>::posv(n, nrhs);
            if (dump) { auto d = to_dense(X); write_raw(a.prefix + ".out.bin", d.data(), d.size()); }
        }
        else {
            std::fprintf(stderr, "posv_mixed needs type d or z\n");
            return 2;
        }
    }
    else if (a.routine == "gesv") {
        // lu_factor + lu_solve_using_factor (test/test_gesv.cc:225-260)
        auto A = make_matrix<T>(n, n, nb, a.seedA, a.get("kind", "rand"));
        auto B = make_matrix<T>(n, nrhs, nb, a.seedB, "rand");
        slate::Pivots pivots;
        auto t0 = tic();
        info = slate::lu_factor(A, pivots, opts);
        // trans=t|c: solve op(A) X = B with the factors of A (src/getrs.cc:97-112; test/test_gesv.cc `trans`)
        const std::string tr = a.get("trans", "n");
        auto opA = A;
        if (tr == "t") opA = slate::transpose(A); else if (tr == "c") opA = slate::conj_transpose(A);
        if (info == 0) slate::lu_solve_using_factor(opA, pivots, B, opts);
        seconds = toc(t0);
        gflop = lapack::Gflop<T>::gesv(n, nrhs);
        if (dump) { auto d = to_dense(B); write_raw(a.prefix + ".out.bin", d.data(), d.size()); }
    }
    else if (a.routine == "hemm") {
        // C = alpha A B + beta C (side=l, default: A n x n, B and C n x nrhs) or C = alpha B A + beta C (side=r: B and C
        // nrhs x n), A Hermitian (lower)  (test/test_hemm.cc:92-194)
        const bool right = a.get("side", "l") == "r";
        const slate::Side side = right ? slate::Side::Right : slate::Side::Left;
        slate::HermitianMatrix<T> A(slate::Uplo::Lower, n, nb, slate::GridOrder::Col, g_p, g_q, MPI_COMM_WORLD);
        A.insertLocalTiles();
        slate::MatgenParams p; p.verbose = 0; p.kind = a.get("kind", "rand"); p.seed = a.seedA;
        p.cond_request = p.cond_actual = p.condD = NAN;
        slate::generate_matrix(p, A);
        auto B = right ? make_matrix<T>(nrhs, n, nb, a.seedB, "rand") : make_matrix<T>(n, nrhs, nb, a.seedB, "rand");
        auto C = right ? make_matrix<T>(nrhs, n, nb, a.seedC, "rand") : make_matrix<T>(n, nrhs, nb, a.seedC, "rand");
        auto t0 = tic();
        slate::hemm(side, alpha, A, B, beta, C, opts);
        seconds = toc(t0);
        gflop = right ? blas::Gflop<T>::hemm(side, nrhs, n) : blas::Gflop<T>::hemm(side, n, nrhs);
        if (dump) { auto d = to_dense(C); write_raw(a.prefix + ".out.bin", d.data(), d.size()); }
    }
    else if (a.routine == "her2k") {
        // C = alpha A B^H + conj(alpha) B A^H + beta C, C Hermitian lower (test/test_her2k.cc; slate::her2k, src/her2k.cc)
        int64_t k = a.geti("k", n);
        const bool tr = a.get("trans", "n") != "n";          // trans=c: A, B stored k x n, conjugate-transposed views
        auto A = tr ? make_matrix<T>(k, n, nb, a.seedA, "rand") : make_matrix<T>(n, k, nb, a.seedA, "rand");
        auto B = tr ? make_matrix<T>(k, n, nb, a.seedB, "rand") : make_matrix<T>(n, k, nb, a.seedB, "rand");
        slate::HermitianMatrix<T> C(slate::Uplo::Lower, n, nb, slate::GridOrder::Col, g_p, g_q, MPI_COMM_WORLD);
        C.insertLocalTiles();
        slate::MatgenParams p; p.verbose = 0; p.kind = "rand"; p.seed = a.seedC;
        p.cond_request = p.cond_actual = p.condD = NAN;
        slate::generate_matrix(p, C);
        real_t rb = std::real(beta);
        auto t0 = tic();
        auto opA = A, opB = B;
        if (tr) { opA = slate::conj_transpose(A); opB = slate::conj_transpose(B); }
        slate::her2k(alpha, opA, opB, rb, C, opts);
        seconds = toc(t0);
        gflop = blas::Gflop<T>::her2k(n, k);
        if (dump) {
            auto d = tz_to_dense<slate::HermitianMatrix<T>, T>(C, true);
            write_raw(a.prefix + ".out.bin", d.data(), d.size());
        }
    }
    else if (a.routine == "syrk" || a.routine == "syr2k") {
        // complex-symmetric rank-k / rank-2k updates (no conjugation; slate::syrk, src/syrk.cc; slate::syr2k, src/syr2k.cc):
        // C = alpha A A^T + beta C   /   C = alpha A B^T + alpha B A^T + beta C,  C symmetric lower, alpha / beta scalar_t
        int64_t k = a.geti("k", n);
        const bool tr = a.get("trans", "n") != "n";          // trans=t: A, B stored k x n, transposed views
        auto A = tr ? make_matrix<T>(k, n, nb, a.seedA, "rand") : make_matrix<T>(n, k, nb, a.seedA, "rand");
        auto B = tr ? make_matrix<T>(k, n, nb, a.seedB, "rand") : make_matrix<T>(n, k, nb, a.seedB, "rand");
        slate::SymmetricMatrix<T> C(slate::Uplo::Lower, n, nb, slate::GridOrder::Col, g_p, g_q, MPI_COMM_WORLD);
        C.insertLocalTiles();
        slate::MatgenParams p; p.verbose = 0; p.kind = "rand"; p.seed = a.seedC;
        p.cond_request = p.cond_actual = p.condD = NAN;
        slate::generate_matrix(p, C);
        auto t0 = tic();
        auto opA = A, opB = B;
        if (tr) { opA = slate::transpose(A); opB = slate::transpose(B); }
        if (a.routine == "syrk") slate::syrk(alpha, opA, beta, C, opts);
        else                     slate::syr2k(alpha, opA, opB, beta, C, opts);
        seconds = toc(t0);
        gflop = a.routine == "syrk" ? blas::Gflop<T>::syrk(n, k) : blas::Gflop<T>::syr2k(n, k);
        if (dump) {
            auto d = tz_to_dense<slate::SymmetricMatrix<T>, T>(C, true);
            write_raw(a.prefix + ".out.bin", d.data(), d.size());
        }
    }
    else if (a.routine == "trmm") {
        // B = alpha op(A) B (side=l, default: A m x m) or B = alpha B op(A) (side=r: A n x n), B m x n, A lower triangular
        // (rand), op=n|t|c as a transposed view of A  (test/test_trmm.cc:84-174; slate::trmm, src/trmm.cc)
        int64_t m = a.geti("m", n);
        bool unit = a.get("diag", "n") == "u";
        const bool right = a.get("side", "l") == "r";
        const slate::Side side = right ? slate::Side::Right : slate::Side::Left;
        const std::string op = a.get("op", "n");
        slate::TriangularMatrix<T> A(slate::Uplo::Lower, unit ? slate::Diag::Unit : slate::Diag::NonUnit, right ? n : m, nb,
                                     slate::GridOrder::Col, g_p, g_q, MPI_COMM_WORLD);
        A.insertLocalTiles();
        slate::MatgenParams p; p.verbose = 0; p.kind = "rand"; p.seed = a.seedA;
        p.cond_request = p.cond_actual = p.condD = NAN;
        slate::generate_matrix(p, A);
        auto B = make_matrix<T>(m, n, nb, a.seedB, "rand");
        auto opA = A;
        if (op == "t") opA = slate::transpose(A);
        else if (op == "c") opA = slate::conj_transpose(A);
        auto t0 = tic();
        slate::trmm(side, alpha, opA, B, opts);
        seconds = toc(t0);
        gflop = blas::Gflop<T>::trmm(side, m, n);
        if (dump) { auto d = to_dense(B); write_raw(a.prefix + ".out.bin", d.data(), d.size()); }
    }
    else if (a.routine == "symm") {
        // C = alpha A B + beta C (side=l) or C = alpha B A + beta C (side=r: B and C nrhs x n), A complex-symmetric (lower)
        // (test/test_symm.cc; slate::symm, src/symm.cc)
        const bool right = a.get("side", "l") == "r";
        const slate::Side side = right ? slate::Side::Right : slate::Side::Left;
        slate::SymmetricMatrix<T> A(slate::Uplo::Lower, n, nb, slate::GridOrder::Col, g_p, g_q, MPI_COMM_WORLD);
        A.insertLocalTiles();
        slate::MatgenParams p; p.verbose = 0; p.kind = a.get("kind", "rand"); p.seed = a.seedA;
        p.cond_request = p.cond_actual = p.condD = NAN;
        slate::generate_matrix(p, A);
        auto B = right ? make_matrix<T>(nrhs, n, nb, a.seedB, "rand") : make_matrix<T>(n, nrhs, nb, a.seedB, "rand");
        auto C = right ? make_matrix<T>(nrhs, n, nb, a.seedC, "rand") : make_matrix<T>(n, nrhs, nb, a.seedC, "rand");
        auto t0 = tic();
        slate::symm(side, alpha, A, B, beta, C, opts);
        seconds = toc(t0);
        gflop = right ? blas::Gflop<T>::symm(side, nrhs, n) : blas::Gflop<T>::symm(side, n, nrhs);
        if (dump) { auto d = to_dense(C); write_raw(a.prefix + ".out.bin", d.data(), d.size()); }
    }
    else if (a.routine == "norms") {
        // max / one / inf / fro of a general rand matrix: slate::norm (src/norm.cc)
        int64_t m = a.geti("m", n);
        auto A = make_matrix<T>(m, n, nb, a.seedA, a.get("kind", "rand"));
        if (dump) { auto d = to_dense(A); write_raw(a.prefix + ".A.bin", d.data(), d.size()); }
        real_t v[4];
        auto t0 = tic();
        v[0] = slate::norm(slate::Norm::Max, A, opts);
        v[1] = slate::norm(slate::Norm::One, A, opts);
        v[2] = slate::norm(slate::Norm::Inf, A, opts);
        v[3] = slate::norm(slate::Norm::Fro, A, opts);
        seconds = toc(t0);
        double vd[4] = { double(v[0]), double(v[1]), double(v[2]), double(v[3]) };
        write_raw(a.prefix + ".out.bin", vd, 4);
    }
    else {
        std::fprintf(stderr, "unknown routine %s\n", a.routine.c_str());
        return 2;
    }
    if (g_rank == 0)
    std::printf("{\"routine\": \"%s\", \"type\": \"%s\", \"n\": %lld, \"nb\": %lld, \"seconds\": %.6f, "
                "\"gflops\": %.3f, \"threads\": %d, \"info\": %lld, \"iters\": %d, \"target\": \"HostTask\"}\n",
                a.routine.c_str(), a.type.c_str(), (long long) n, (long long) nb, seconds,
                seconds > 0 ? gflop / seconds : 0.0, omp_get_max_threads(), (long long) info, iters);
    return 0;
}

} // namespace

int main(int argc, char** argv)
{
    if (argc < 9) {
        std::fprintf(stderr,
            "usage: %s ROUTINE TYPE n nb seedA seedB seedC OUTPREFIX [key=value ...]\n", argv[0]);
        return 2;
    }
    int provided = 0;
    MPI_Init_thread(&argc, &argv, MPI_THREAD_MULTIPLE, &provided);
    Args a;
    a.routine = argv[1]; a.type = argv[2];
    a.n = std::atoll(argv[3]); a.nb = std::atoll(argv[4]);
    a.seedA = std::atoll(argv[5]); a.seedB = std::atoll(argv[6]); a.seedC = std::atoll(argv[7]);
    a.prefix = argv[8];
    for (int i = 9; i < argc; ++i) {
        std::string s = argv[i];
        auto eq = s.find('=');
        if (eq != std::string::npos) a.kv[s.substr(0, eq)] = s.substr(eq + 1);
    }
    if (a.kv.count("threads")) omp_set_num_threads(std::atoi(a.kv["threads"].c_str()));
    MPI_Comm_rank(MPI_COMM_WORLD, &g_rank);
    MPI_Comm_size(MPI_COMM_WORLD, &g_world);
    g_p = int(a.geti("p", 1)); g_q = int(a.geti("q", 1));
    if (g_p * g_q != g_world) {
        std::fprintf(stderr, "ref_dump: grid %d x %d needs %d ranks, the world has %d\n", g_p, g_q, g_p * g_q, g_world);
        return 2;
    }
    if (g_world > 1) a.prefix += ".r" + std::to_string(g_rank);
    int rc = 2;
    if      (a.type == "d") rc = run<double>(a);
    else if (a.type == "z") rc = run<std::complex<double>>(a);
    else if (a.type == "s") rc = run<float>(a);
    else if (a.type == "c") rc = run<std::complex<float>>(a);
    MPI_Finalize();
    return rc;
}
