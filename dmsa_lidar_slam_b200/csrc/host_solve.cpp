// host_solve.cpp — dense P x P solvers of the LM step (DmsaOptimizer.h:107-113), host side.
// Compiled by the host C++ compiler (not nvcc) so that GCC function multi-versioning can emit AVX-512 / AVX2 / baseline
// clones of the same loops; all arithmetic is element-wise IEEE double without FMA contraction (-ffp-contract=off),
// so every clone produces bit-identical results.
#include <atomic>
#include <cmath>
#include <condition_variable>
#include <cstddef>
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <mutex>
#include <thread>
#include <utility>
#include <vector>

#if defined(__x86_64__) && defined(__GNUC__) && !defined(__clang__)
#define DMSA_CLONES __attribute__((target_clones("avx512f", "avx2", "default")))
#else
#define DMSA_CLONES
#endif

// Eigen dynamic inverse() == PartialPivLU: explicit inverse by LU with partial pivoting (DmsaOptimizer.h:113).
// All n right-hand sides are substituted together, row by row (contiguous axpy loops the host compiler vectorises);
// every element still sees exactly the operation sequence of a column-by-column substitution (j ascending).
// The elimination is organised in panels of LU_NB pivots: inside a panel the steps run as in the textbook (pivot search,
// row exchange over the whole row, multipliers, update) but touch the panel's columns only; the columns to the right of
// the panel then take all of the panel's steps at once, row by row, with a row's columns held in registers while the
// pivot rows stream past.  Every element still receives a_ij -= f_ik a_kj for ascending k with the very same operands
// (the multipliers travel with their rows through the exchanges), so the factors equal the unblocked elimination's bit
// for bit — and the device kernels' (kernels_solve.cuh) — while the matrix is read and written once per panel, not once
// per step.
#define LU_NB 16
template <int CB>
static inline __attribute__((always_inline)) void lu_trailing_cols(double* __restrict__ a, int n, int i, int k0, int kend, int c) {
    double* ai = a + (size_t)i * n;
    double acc[CB];
    for (int q = 0; q < CB; ++q) acc[q] = ai[c + q];
    for (int k = k0; k < kend; ++k) {
        const double f = ai[k];
        const double* ak = a + (size_t)k * n + c;
        for (int q = 0; q < CB; ++q) acc[q] -= f * ak[q];
    }
    for (int q = 0; q < CB; ++q) ai[c + q] = acc[q];
}
DMSA_CLONES static void lu_factor_impl(std::vector<double>& av, std::vector<int>& piv, int n) {
    double* a = av.data();
    for (int i = 0; i < n; ++i) piv[i] = i;
    for (int k0 = 0; k0 < n; k0 += LU_NB) {
        const int k1 = std::min(n, k0 + LU_NB);
        for (int k = k0; k < k1; ++k) {  // the panel's steps on the panel's columns
            int p = k;
            double best = std::fabs(a[(size_t)k * n + k]);
            for (int i = k + 1; i < n; ++i) {
                double v = std::fabs(a[(size_t)i * n + k]);
                if (v > best) {
                    best = v;
                    p = i;
                }
            }
            if (p != k) {
                for (int j = 0; j < n; ++j) std::swap(a[(size_t)k * n + j], a[(size_t)p * n + j]);
                std::swap(piv[k], piv[p]);
            }
            const double d = a[(size_t)k * n + k];
            const double* ak = a + (size_t)k * n;
            for (int i = k + 1; i < n; ++i) {
                double* ai = a + (size_t)i * n;
                const double f = ai[k] / d;
                ai[k] = f;
                for (int j = k + 1; j < k1; ++j) ai[j] -= f * ak[j];
            }
        }
        for (int i = k0 + 1; i < n; ++i) {  // the columns right of the panel: row i takes the steps of the pivots above it
            const int kend = std::min(i, k1);
            int c = k1;
            for (; c + 32 <= n; c += 32) lu_trailing_cols<32>(a, n, i, k0, kend, c);
            for (; c + 8 <= n; c += 8) lu_trailing_cols<8>(a, n, i, k0, kend, c);
            for (; c < n; ++c) lu_trailing_cols<1>(a, n, i, k0, kend, c);
        }
    }
}
// forward + back substitution of the right-hand-side columns [c0, c1) (independent of every other column).
// Forward: x_i -= l_ij x_j with ascending j.  Backward, column-oriented: x_j *= 1 / u_jj once every row below has been
// applied, then x_i -= u_ij x_j for the rows above, descending j (Eigen's triangular matrix solver, the path inverse() takes,
// also scales by the reciprocal diagonal; neither order can be pinned against Eigen here).  This is the order in which
// a column's dependency chain is 2 n steps long, so the device solver (kernels_solve.cuh) runs the same sequence with one
// warp per column.
// Both sweeps run row by row with the CB columns of a row held in registers while the row's whole dependency list streams
// past (one load of x_j per multiply-subtract instead of a load and a store of x_i as well); an element still sees
// exactly the sequence described above — forward: j ascending; backward: j descending, then the reciprocal diagonal.
template <int CB>
static inline __attribute__((always_inline)) void lu_subst_cols(const double* __restrict__ a, double* __restrict__ inv, int n, int ldx, int c) {
    for (int i = 0; i < n; ++i) {  // forward substitution, unit lower triangle
        double* xi = inv + (size_t)i * ldx + c;
        const double* ai = a + (size_t)i * n;
        double acc[CB];
        for (int q = 0; q < CB; ++q) acc[q] = xi[q];
        for (int j = 0; j < i; ++j) {
            const double l = ai[j];
            const double* xj = inv + (size_t)j * ldx + c;
            for (int q = 0; q < CB; ++q) acc[q] -= l * xj[q];
        }
        for (int q = 0; q < CB; ++q) xi[q] = acc[q];
    }
    for (int i = n - 1; i >= 0; --i) {  // back substitution
        double* xi = inv + (size_t)i * ldx + c;
        const double* ai = a + (size_t)i * n;
        double acc[CB];
        for (int q = 0; q < CB; ++q) acc[q] = xi[q];
        for (int j = n - 1; j > i; --j) {
            const double u = ai[j];
            const double* xj = inv + (size_t)j * ldx + c;
            for (int q = 0; q < CB; ++q) acc[q] -= u * xj[q];
        }
        const double rdiag = 1.0 / ai[i];
        for (int q = 0; q < CB; ++q) xi[q] = acc[q] * rdiag;
    }
}
DMSA_CLONES static void lu_subst_block_impl(const double* a, double* inv, int n, int ldx, int c0, int c1) {
    if (c0 >= c1) return;
    if (c1 == n) c1 = std::min(ldx, (n + 7) / 8 * 8);  // the block that ends the matrix also takes the (zero) padding columns of its last line
    int c = c0;
    for (; c + 32 <= c1; c += 32) lu_subst_cols<32>(a, inv, n, ldx, c);
    for (; c + 8 <= c1; c += 8) lu_subst_cols<8>(a, inv, n, ldx, c);
    for (; c < c1; ++c) lu_subst_cols<1>(a, inv, n, ldx, c);
}

// A few helper threads for the substitution of the n independent right-hand sides.  The optimizer loop knows when the
// solve is about to happen (right after the read-back of [H | g]), so it ARMS the pool before it blocks on the stream:
// armed workers spin on the job counter and start without a wake-up latency; disarmed workers sleep on a condition
// variable.  Each column is still computed by exactly one thread with the serial operation order: results are identical.
namespace {
// columns per thread: a multiple of 8 doubles (one 64-byte cache line) so that no two threads ever write the same line of a row
inline int colBlock(int n) { return ((n + 3) / 4 + 7) / 8 * 8; }
struct SolvePool {
    static constexpr int kWorkers = 3;
    std::vector<std::thread> th;
    std::mutex m;
    std::condition_variable cv;
    std::atomic<int> armed{0};
    std::atomic<unsigned long long> gen{0};
    std::atomic<int> remaining{0};
    std::atomic_flag busy = ATOMIC_FLAG_INIT;  // one solve at a time uses the helpers; a concurrent one runs serially
    const double* a = nullptr;
    double* inv = nullptr;
    int n = 0, ldx = 0;
    bool stop = false;
    bool disabled = false;
    void start() {
        std::lock_guard<std::mutex> lk(m);  // two contexts may arm from two host threads at the same time
        if (!th.empty() || disabled) return;
        // Three helper threads take cache-line aligned column blocks of the inverse while the caller takes the first
        // (measured on the B200 host, P = 114: 0.25 ms alone, 0.19 ms with helpers).  DMSA_B200_SOLVER_THREADS=0 turns
        // them off, =1 forces them on; by default they need 8 hardware threads PER LOCAL RANK (one process per GPU under
        // torchrun: LOCAL_WORLD_SIZE), so that N ranks on a small host do not oversubscribe it with spinning helpers.
        const char* e = std::getenv("DMSA_B200_SOLVER_THREADS");
        const char* lws = std::getenv("LOCAL_WORLD_SIZE");
        const unsigned ranks = (lws && std::atoi(lws) > 0) ? (unsigned)std::atoi(lws) : 1u;
        if ((e && std::atoi(e) <= 0) || (!e && std::thread::hardware_concurrency() / ranks < 8)) {
            disabled = true;
            return;
        }
        for (int w = 0; w < kWorkers; ++w) th.emplace_back([this, w] { worker(w); });
    }
    void worker(int w) {
        unsigned long long seen = 0;  // generations start at 0; a job published before this thread first runs must not be missed
        for (;;) {
            {
                std::unique_lock<std::mutex> lk(m);
                cv.wait(lk, [&] { return stop || armed.load() > 0 || gen.load() != seen; });
                if (stop) return;
            }
            // `armed` is a reference count: every arm() holds one reference and the publisher of a job holds one for the
            // job's duration, so a published generation is always consumed (the count cannot reach 0 while a job is open)
            while (armed.load(std::memory_order_acquire) > 0) {
                const unsigned long long g = gen.load(std::memory_order_acquire);
                if (g != seen) {
                    seen = g;
                    const int blk = w + 1;
                    lu_subst_block_impl(a, inv, n, ldx, std::min(n, colBlock(n) * blk), std::min(n, colBlock(n) * (blk + 1)));
                    remaining.fetch_sub(1, std::memory_order_acq_rel);
                } else {
                    __builtin_ia32_pause();
                }
            }
        }
    }
    ~SolvePool() {
        {
            std::lock_guard<std::mutex> lk(m);
            stop = true;
            armed.store(0);
        }
        cv.notify_all();
        for (auto& t : th) t.join();
    }
};
SolvePool g_pool;
}  // namespace

static void pool_ref() {
    {
        std::lock_guard<std::mutex> lk(g_pool.m);
        g_pool.armed.fetch_add(1, std::memory_order_acq_rel);
    }
    g_pool.cv.notify_all();
}
static void pool_unref() { g_pool.armed.fetch_sub(1, std::memory_order_acq_rel); }
// arm / disarm are reference counted: two contexts solving on two host threads cannot switch each other's helpers off
void dmsa_host_solver_arm() {
    g_pool.start();
    pool_ref();
}
void dmsa_host_solver_disarm() { pool_unref(); }

// LU + substitution of P*I into a cache-line aligned, padded block x (n x ldx, ldx a multiple of 8 doubles: the helper
// threads own whole lines of every row).  a: n x n row-major, factorised in place.  Scratch lives in thread-local
// buffers that are reused from call to call.
static double* lu_inverse_padded(std::vector<double>& a, int n, int& ldx_out) {
    thread_local std::vector<int> piv;
    thread_local std::vector<double> xbuf;
    piv.resize(n);
    lu_factor_impl(a, piv, n);
    const int ldx = (n + 7) / 8 * 8;
    xbuf.assign((size_t)n * ldx + 8, 0.0);
    double* x = xbuf.data();
    while (reinterpret_cast<uintptr_t>(x) & 63) ++x;
    for (int i = 0; i < n; ++i) x[(size_t)i * ldx + piv[i]] = 1.0;  // P * I
    if (n >= 64 && g_pool.armed.load(std::memory_order_acquire) > 0 && !g_pool.th.empty() && !g_pool.busy.test_and_set(std::memory_order_acquire)) {
        g_pool.a = a.data();
        g_pool.inv = x;
        g_pool.n = n;
        g_pool.ldx = ldx;
        g_pool.remaining.store(SolvePool::kWorkers, std::memory_order_release);
        pool_ref();  // the publisher's own reference: the helpers stay in their spin loop until this job is consumed
        g_pool.gen.fetch_add(1, std::memory_order_acq_rel);
        lu_subst_block_impl(a.data(), x, n, ldx, 0, std::min(n, colBlock(n)));
        while (g_pool.remaining.load(std::memory_order_acquire) > 0) __builtin_ia32_pause();
        pool_unref();
        g_pool.busy.clear(std::memory_order_release);
    } else {
        lu_subst_block_impl(a.data(), x, n, ldx, 0, n);
    }
    ldx_out = ldx;
    return x;
}
static bool lu_solve_inverse_impl(const std::vector<double>& A, int n, std::vector<double>& inv) {
    std::vector<double> a(A);
    int ldx = 0;
    const double* x = lu_inverse_padded(a, n, ldx);
    inv.resize((size_t)n * n);
    for (int i = 0; i < n; ++i) std::copy(x + (size_t)i * ldx, x + (size_t)i * ldx + n, inv.begin() + (size_t)i * n);
    return true;
}
// The reference's LM step in one go (DmsaOptimizer.h:108-113): H.diag += lambda, step = (-alpha * H.inverse()) * g.
// Same arithmetic as lu_solve_inverse_impl + the caller's product, without the intermediate copies.
bool dmsa_host_lm_step(const double* hg, int n, double lambda, double alpha, double* step) {
    thread_local std::vector<double> a;
    a.assign(hg, hg + (size_t)n * n);
    for (int i = 0; i < n; ++i) a[(size_t)i * n + i] += lambda;
    const double* g = hg + (size_t)n * n;
    int ldx = 0;
    const double* x = lu_inverse_padded(a, n, ldx);
    bool nan = false;
    for (int r = 0; r < n; ++r) {
        const double* xr = x + (size_t)r * ldx;
        double s = 0;
        for (int b = 0; b < n; ++b) s += (-alpha * xr[b]) * g[b];
        step[r] = s;
        if (std::isnan(s)) nan = true;
    }
    return nan;
}

DMSA_CLONES bool lu_solve_vec_impl(const std::vector<double>& A, int n, const double* b, std::vector<double>& x) {
    std::vector<double> a(A);
    x.assign(b, b + n);
    for (int k = 0; k < n; ++k) {
        int p = k;
        double best = std::fabs(a[(size_t)k * n + k]);
        for (int i = k + 1; i < n; ++i) {
            double v = std::fabs(a[(size_t)i * n + k]);
            if (v > best) {
                best = v;
                p = i;
            }
        }
        if (p != k) {
            for (int j = 0; j < n; ++j) std::swap(a[(size_t)k * n + j], a[(size_t)p * n + j]);
            std::swap(x[k], x[p]);
        }
        const double d = a[(size_t)k * n + k];
        const double* __restrict__ ak = &a[(size_t)k * n];
        for (int i = k + 1; i < n; ++i) {
            double* __restrict__ ai = &a[(size_t)i * n];
            const double f = ai[k] / d;
            for (int j = k + 1; j < n; ++j) ai[j] -= f * ak[j];
            x[i] -= f * x[k];
        }
    }
    for (int i = n - 1; i >= 0; --i) {
        double s = x[i];
        const double* ai = &a[(size_t)i * n];
        for (int j = i + 1; j < n; ++j) s -= ai[j] * x[j];
        x[i] = s / ai[i];
    }
    return true;
}

// Cholesky (right-looking, row-major lower triangle) solve of the SPD system (J^T J + lambda I) x = b; false if a pivot
// is not positive (then the caller falls back to LU).  Used by the keyframe-bundle extension only.
DMSA_CLONES bool chol_solve_vec_impl(const std::vector<double>& A, int n, const double* b, std::vector<double>& x) {
    std::vector<double> a(A), col(n);
    for (int k = 0; k < n; ++k) {
        const double d = a[(size_t)k * n + k];
        if (!(d > 0.0)) return false;
        const double lkk = std::sqrt(d);
        a[(size_t)k * n + k] = lkk;
        for (int i = k + 1; i < n; ++i) {
            a[(size_t)i * n + k] /= lkk;
            col[i] = a[(size_t)i * n + k];
        }
        for (int i = k + 1; i < n; ++i) {
            double* __restrict__ ai = &a[(size_t)i * n];
            const double lik = col[i];
            const double* __restrict__ c = col.data();
            for (int j = k + 1; j <= i; ++j) ai[j] -= lik * c[j];
        }
    }
    x.assign(b, b + n);
    for (int i = 0; i < n; ++i) {  // L y = b
        double s = x[i];
        const double* ai = &a[(size_t)i * n];
        for (int j = 0; j < i; ++j) s -= ai[j] * x[j];
        x[i] = s / ai[i];
    }
    for (int i = n - 1; i >= 0; --i) {  // L^T x = y
        double s = x[i];
        for (int j = i + 1; j < n; ++j) s -= a[(size_t)j * n + i] * x[j];
        x[i] = s / a[(size_t)i * n + i];
    }
    return true;
}


// un-cloned entry points (the anonymous-namespace declarations in dmsa_b200.cu bind to these through the linker)
namespace {
}
bool dmsa_host_lu_inverse(const std::vector<double>& A, int n, std::vector<double>& inv) { return lu_solve_inverse_impl(A, n, inv); }
bool dmsa_host_lu_solve(const std::vector<double>& A, int n, const double* b, std::vector<double>& x) { return lu_solve_vec_impl(A, n, b, x); }
bool dmsa_host_chol_solve(const std::vector<double>& A, int n, const double* b, std::vector<double>& x) { return chol_solve_vec_impl(A, n, b, x); }
