// host_solve.cpp — dense P x P solvers of the LM step (DmsaOptimizer.h:107-113), host side.
// Compiled by the host C++ compiler (not nvcc) so that GCC function multi-versioning can emit AVX-512 / AVX2 / baseline
// clones of the same loops; all arithmetic is element-wise IEEE double without FMA contraction (-ffp-contract=off),
// so every clone produces bit-identical results.
#include <cmath>
#include <cstddef>
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
DMSA_CLONES bool lu_solve_inverse_impl(const std::vector<double>& A, int n, std::vector<double>& inv) {
    std::vector<double> a(A);
    std::vector<int> piv(n);
    for (int i = 0; i < n; ++i) piv[i] = i;
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
            std::swap(piv[k], piv[p]);
        }
        const double d = a[(size_t)k * n + k];
        const double* __restrict__ ak = &a[(size_t)k * n];
        for (int i = k + 1; i < n; ++i) {
            double* __restrict__ ai = &a[(size_t)i * n];
            const double f = ai[k] / d;
            ai[k] = f;
            for (int j = k + 1; j < n; ++j) ai[j] -= f * ak[j];
        }
    }
    inv.assign((size_t)n * n, 0.0);
    for (int i = 0; i < n; ++i) inv[(size_t)i * n + piv[i]] = 1.0;  // P * I
    // (measured: an OpenMP team costs more than it saves at P = 114; the blocks stay a plain loop)
    const int nblk = 1;
    for (int blk = 0; blk < nblk; ++blk) {
        const int c0 = (int)((long long)n * blk / nblk), c1 = (int)((long long)n * (blk + 1) / nblk);
        for (int i = 0; i < n; ++i) {  // forward substitution, unit lower triangle
            double* __restrict__ xi = &inv[(size_t)i * n];
            const double* ai = &a[(size_t)i * n];
            for (int j = 0; j < i; ++j) {
                const double l = ai[j];
                const double* __restrict__ xj = &inv[(size_t)j * n];
                for (int c = c0; c < c1; ++c) xi[c] -= l * xj[c];
            }
        }
        for (int i = n - 1; i >= 0; --i) {  // back substitution
            double* __restrict__ xi = &inv[(size_t)i * n];
            const double* ai = &a[(size_t)i * n];
            for (int j = i + 1; j < n; ++j) {
                const double u = ai[j];
                const double* __restrict__ xj = &inv[(size_t)j * n];
                for (int c = c0; c < c1; ++c) xi[c] -= u * xj[c];
            }
            const double dinv = ai[i];
            for (int c = c0; c < c1; ++c) xi[c] = xi[c] / dinv;
        }
    }
    return true;
}

// LU with partial pivoting and ONE right-hand side (no explicit inverse): used by the keyframe-bundle extension, where no
// reference arithmetic exists to mirror and the P x P system is large (P = 378: 3x less work than forming H^-1).
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
