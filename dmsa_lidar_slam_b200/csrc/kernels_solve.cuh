// kernels_solve.cuh — the Levenberg-Marquardt step of DmsaOptimizer.h:107-128 on the device, so that an iteration
// does not have to stop for a host round trip between the Jacobian pass and the line search:
//
//   H.diagonal().array() += lambda_diag                                   :108-111
//   step = (-alpha * H.inverse()) * (J^T e0)                              :113  (Eigen dynamic inverse() == PartialPivLU)
//   NaN guard                                                             :115-122
//   if (max|step| > max_step) step = (max_step / max|step|) * step        :124-128
//
// One thread block.  The arithmetic is the operation sequence of host_solve.cpp (lu_factor_impl / lu_subst_block_impl /
// solveStep), element for element: row-major LU with partial pivoting (first maximum wins), explicit inverse by
// substituting the n columns of P*I (forward with ascending j, backward with ascending j, division last), then the
// matrix-vector product accumulated in ascending column order.  IEEE double, no FMA contraction (the library is built
// with -fmad=false), so host and device produce bit-identical steps (tests/test_gpu_parity.py::test_device_lm_solve...).
// The matrices live in shared memory when both fit (P <= 116: 2 x 104 KB), else in the caller's global scratch (L2).
#pragma once
#include <cuda_runtime.h>

namespace dmsa {

#define LM_SOLVE_T 1024

struct LmSolveArgs {
    const double* hg;  // [H (P*P row-major) | g (P) | err0]
    int P;
    int lda;           // leading dimension (odd: fewer bank conflicts on column walks)
    double lambda, alpha, max_step;
    double* scratch;   // global scratch 2 * P * lda doubles (used when the matrices do not fit in shared memory)
    int use_smem;
    double* step;      // out [P] (input of the line-search batch)
    double* step2;     // out [P] second copy inside the iteration's read-back block (may be null)
    long long* clk;    // optional debug: cycle stamps of the phases (thread 0), 6 entries
    double* tail;      // out [0] = err0 (copied from hg), [1] = 1.0 if the step contains NaN (then left unclamped), else 0.0
};

__global__ void __launch_bounds__(LM_SOLVE_T, 1) k_lm_solve(LmSolveArgs q) {
    extern __shared__ __align__(16) unsigned char lm_smem_raw[];
    __shared__ double wv[LM_SOLVE_T / 32], ws[LM_SOLVE_T / 32];
    __shared__ int wi[LM_SOLVE_T / 32];
    __shared__ double s_red[2][LM_SOLVE_T / 32];
    __shared__ int s_nan;
    const int n = q.P, lda = q.lda, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    double* __restrict__ A = q.use_smem ? reinterpret_cast<double*>(lm_smem_raw) : q.scratch;
    double* __restrict__ X = A + (size_t)n * lda;
    int* __restrict__ piv = reinterpret_cast<int*>(q.use_smem ? (void*)(X + (size_t)n * lda) : (void*)(q.scratch + 2 * (size_t)n * lda));
    // load H + lambda I, X = 0
    for (int e = tid; e < n * n; e += LM_SOLVE_T) {
        const int i = e / n, j = e - i * n;
        double v = q.hg[e];
        if (i == j) v += q.lambda;
        A[(size_t)i * lda + j] = v;
        X[(size_t)i * lda + j] = 0.0;
    }
    for (int i = tid; i < n; i += LM_SOLVE_T) piv[i] = i;
    if (tid == 0) s_nan = 0;
    __syncthreads();
    if (q.clk && tid == 0) q.clk[0] = clock64();
    // ---- LU with partial pivoting (host_solve.cpp lu_factor_impl) ----
    for (int k = 0; k < n; ++k) {
        // pivot: first maximum of |a[i][k]|, i >= k; a NaN never replaces the running best (v > best is false), and a NaN
        // at i == k is never replaced.  The signed pivot value travels with the key.
        double v = -1.0, sv = 0.0;
        int idx = 0x7fffffff;
        if (tid < n - k) {
            idx = k + tid;
            sv = A[(size_t)idx * lda + k];
            v = fabs(sv);
            if (v != v) v = (idx == k) ? __longlong_as_double(0x7ff0000000000000ll) : -1.0;
        }
        const int nw = (n - k + 31) >> 5;  // warps holding candidates
        if (wid < nw) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double v2 = __shfl_down_sync(0xffffffffu, v, o);
                const double s2 = __shfl_down_sync(0xffffffffu, sv, o);
                const int i2 = __shfl_down_sync(0xffffffffu, idx, o);
                if (v2 > v || (v2 == v && i2 < idx)) {
                    v = v2;
                    sv = s2;
                    idx = i2;
                }
            }
            if (lane == 0) {
                wv[wid] = v;
                ws[wid] = sv;
                wi[wid] = idx;
            }
        }
        __syncthreads();
        double bv = wv[0], d = ws[0];
        int p = wi[0];
        for (int w = 1; w < nw; ++w) {
            const double v2 = wv[w];
            const int i2 = wi[w];
            if (v2 > bv || (v2 == bv && i2 < p)) {
                bv = v2;
                d = ws[w];
                p = i2;
            }
        }
        // row swap k <-> p (all columns but k) and, in the same phase, column k: the multipliers f_i = a[i][k] / d of the
        // rows below the diagonal (row p receives the old a[k][k]); only the thread of row i touches column k of rows k, p
        if (p != k) {  // block-uniform
            for (int j = tid; j < n; j += LM_SOLVE_T) {
                if (j == k) continue;
                const double t = A[(size_t)k * lda + j];
                A[(size_t)k * lda + j] = A[(size_t)p * lda + j];
                A[(size_t)p * lda + j] = t;
            }
            if (tid == LM_SOLVE_T - 1) {
                const int t = piv[k];
                piv[k] = piv[p];
                piv[p] = t;
            }
        }
        if (tid < n - k - 1) {
            const int i = k + 1 + tid;
            double num;
            if (i == p) {
                num = A[(size_t)k * lda + k];
                A[(size_t)k * lda + k] = d;
            } else {
                num = A[(size_t)i * lda + k];
            }
            A[(size_t)i * lda + k] = num / d;
        }
        __syncthreads();
        for (int i = k + 1 + wid; i < n; i += LM_SOLVE_T / 32) {
            const double f = A[(size_t)i * lda + k];
            for (int j = k + 1 + lane; j < n; j += 32) A[(size_t)i * lda + j] = A[(size_t)i * lda + j] - f * A[(size_t)k * lda + j];
        }
        __syncthreads();
    }
    if (q.clk && tid == 0) q.clk[1] = clock64();
    // ---- explicit inverse: substitution of the columns of P*I (lu_subst_block_impl) ----
    for (int i = tid; i < n; i += LM_SOLVE_T) X[(size_t)i * lda + piv[i]] = 1.0;
    __syncthreads();
    // forward substitution (unit lower triangle), right-looking: once row j is final every row below takes its term
    // x[i][c] -= l[i][j] * x[j][c].  Element (i, c) sees j = 0, 1, .., i-1 in ascending order, exactly like the host loop.
    for (int j = 0; j < n - 1; ++j) {
        const int rows = n - j - 1;
        for (int e = tid; e < rows * n; e += LM_SOLVE_T) {
            const int r = e / n, c = e - r * n, i = j + 1 + r;
            X[(size_t)i * lda + c] = X[(size_t)i * lda + c] - A[(size_t)i * lda + j] * X[(size_t)j * lda + c];
        }
        __syncthreads();
    }
    if (q.clk && tid == 0) q.clk[2] = clock64();
    // back substitution: element (i, c) subtracts its terms in ASCENDING j = i+1 .. n-1 (host order) and divides last, so
    // it cannot start before x[i+1] is final: one thread per column walks the rows upwards; the products are formed
    // eight at a time so that the dependent chain is the subtractions only
    if (tid < n) {
        const int c = tid;
        for (int i = n - 1; i >= 0; --i) {
            double xi = X[(size_t)i * lda + c];
            const double* __restrict__ ai = A + (size_t)i * lda;
            int j = i + 1;
            for (; j + 8 <= n; j += 8) {
                double pr[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) pr[u] = ai[j + u] * X[(size_t)(j + u) * lda + c];
#pragma unroll
                for (int u = 0; u < 8; ++u) xi = xi - pr[u];
            }
            for (; j < n; ++j) xi = xi - ai[j] * X[(size_t)j * lda + c];
            X[(size_t)i * lda + c] = xi / ai[i];
        }
    }
    __syncthreads();
    if (q.clk && tid == 0) q.clk[3] = clock64();
    // ---- step = (-alpha * Hinv) * g, NaN guard, infinity-norm clamp (dmsa_b200.cu solveStep) ----
    const double* __restrict__ g = q.hg + (size_t)n * n;
    double s = 0.0;
    double mx = -__longlong_as_double(0x7ff0000000000000ll), mn = __longlong_as_double(0x7ff0000000000000ll);
    if (tid < n) {
        const double na = -q.alpha;
        for (int b = 0; b < n; ++b) s += (na * X[(size_t)tid * lda + b]) * g[b];
        if (s != s) s_nan = 1;
        mx = s;  // std::max(mx, v) / std::min(mn, v): a NaN never wins, and the NaN path returns before the clamp
        mn = s;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    }
    if (lane == 0) {
        s_red[0][wid] = mx;
        s_red[1][wid] = mn;
    }
    __syncthreads();
    if (q.clk && tid == 0) q.clk[4] = clock64();
    if (tid == 0) q.tail[0] = q.hg[(size_t)n * n + n];
    if (s_nan) {
        if (tid < n) {
            q.step[tid] = s;
            if (q.step2) q.step2[tid] = s;
        }
        if (tid == 0) q.tail[1] = 1.0;
        return;
    }
    for (int w = 0; w < LM_SOLVE_T / 32; ++w) {
        mx = fmax(mx, s_red[0][w]);
        mn = fmin(mn, s_red[1][w]);
    }
    const double maxElem = fmax(mx, -mn);
    if (tid < n) {
        if (maxElem > q.max_step) s = (q.max_step / maxElem) * s;
        q.step[tid] = s;
        if (q.step2) q.step2[tid] = s;
    }
    if (tid == 0) q.tail[1] = 0.0;
}


// =====================================================================================================================
// Fast path for P <= 128 (the sliding-window pass: P = 6 * (Nposes - 1) = 114 at 20 control poses), three launches:
//   k_lu_tile   one block, the matrix lives in REGISTERS (4 x 4 elements per thread, rows on lanes, column groups on warps):
//               the pivot search and the multipliers of column k are one warp's business (shuffles), the row swap is a
//               lane-to-lane exchange inside every warp, the rank-1 update is 32 register DMUL/DADD with no memory traffic;
//               ONE barrier per elimination step (the multiplier column is double-buffered in shared memory).
//   k_inv_cols  the n columns of the inverse are independent: one warp per 32 columns (lane = column), L\U staged in
//               shared memory, forward and back substitution in the host's order (ascending j, division last).
//   k_step_fin  step = (-alpha * Hinv) * g in ascending column order, NaN guard, infinity-norm clamp.
// Same operation sequence per element as host_solve.cpp -> bit-identical steps.
// =====================================================================================================================
#define LU_TILE_N 128
struct LuTileArgs {
    const double* hg;  // [H | g | err0]
    int n, lda;
    double lambda;
    double* LU;        // out: n x lda row-major, L (unit, below the diagonal) and U
    int* piv;          // out: row permutation, piv[i] = original row now at position i
};

#define LU_TILE_T 512                 // 16 warps: warp w holds the columns j = w + 16 c (c < 8), lane l the rows i = l + 32 r (r < 4)
#define LU_TILE_W (LU_TILE_T / 32)
#define LU_PICKROW(c, s) ((s) == 0 ? a[0][c] : (s) == 1 ? a[1][c] : (s) == 2 ? a[2][c] : a[3][c])

// Column step of the warp that holds column k (compile-time column slot CK): pivot search, multipliers, bookkeeping.
#define LU_COLSTEP(CK)                                                                                                  \
    {                                                                                                                   \
        double colv[4] = {a[0][CK], a[1][CK], a[2][CK], a[3][CK]};                                                      \
        double bv = -1.0, bs = 0.0;                                                                                     \
        int bi = 0x7fffffff;                                                                                            \
        _Pragma("unroll") for (int r = 0; r < 4; ++r) {                                                                 \
            const int i = lane + 32 * r;                                                                                \
            if (i >= k && i < n) {                                                                                      \
                const double sv = colv[r];                                                                              \
                double v = fabs(sv);                                                                                    \
                if (v != v) v = (i == k) ? INF : -1.0;                                                                  \
                if (v > bv) {                                                                                           \
                    bv = v;                                                                                             \
                    bs = sv;                                                                                            \
                    bi = i;                                                                                             \
                }                                                                                                       \
            }                                                                                                           \
        }                                                                                                               \
        _Pragma("unroll") for (int o = 16; o > 0; o >>= 1) {                                                            \
            const double v2 = __shfl_xor_sync(0xffffffffu, bv, o);                                                      \
            const double s2 = __shfl_xor_sync(0xffffffffu, bs, o);                                                      \
            const int i2 = __shfl_xor_sync(0xffffffffu, bi, o);                                                         \
            if (v2 > bv || (v2 == bv && i2 < bi)) {                                                                     \
                bv = v2;                                                                                                \
                bs = s2;                                                                                                \
                bi = i2;                                                                                                \
            }                                                                                                           \
        }                                                                                                               \
        const int p = bi;                                                                                               \
        const double d = bs;                                                                                            \
        const double mine = rk == 0 ? colv[0] : rk == 1 ? colv[1] : rk == 2 ? colv[2] : colv[3];                        \
        const double akk = __shfl_sync(0xffffffffu, mine, lk); /* a[k][k] before the swap */                            \
        _Pragma("unroll") for (int r = 0; r < 4; ++r) {                                                                 \
            const int i = lane + 32 * r;                                                                                \
            if (i > k && i < n) {                                                                                       \
                const double num = (i == p) ? akk : colv[r];                                                            \
                const double f = num / d;                                                                               \
                colv[r] = f;                                                                                            \
                lcol[buf][i] = f;                                                                                       \
            } else if (i == k) {                                                                                        \
                colv[r] = d;                                                                                            \
            }                                                                                                           \
        }                                                                                                               \
        a[0][CK] = colv[0];                                                                                             \
        a[1][CK] = colv[1];                                                                                             \
        a[2][CK] = colv[2];                                                                                             \
        a[3][CK] = colv[3];                                                                                             \
        if (lane == 0) {                                                                                                \
            s_p[buf] = p;                                                                                               \
            const int t = s_piv[k];                                                                                     \
            s_piv[k] = s_piv[p];                                                                                        \
            s_piv[p] = t;                                                                                               \
        }                                                                                                               \
    }

__global__ void __launch_bounds__(LU_TILE_T, 1) k_lu_tile(LuTileArgs q) {
    __shared__ double lcol[2][LU_TILE_N];
    __shared__ int s_p[2];
    __shared__ int s_piv[LU_TILE_N];
    const int n = q.n, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const double INF = __longlong_as_double(0x7ff0000000000000ll);
    double a[4][8];  // a[r][c] = A[lane + 32 r][warp + 16 c]
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const int i = lane + 32 * r, j = warp + LU_TILE_W * c;
            double v = 0.0;
            if (i < n && j < n) {
                v = __ldg(q.hg + (size_t)i * n + j);
                if (i == j) v += q.lambda;
            }
            a[r][c] = v;
        }
    if (threadIdx.x < LU_TILE_N) s_piv[threadIdx.x] = threadIdx.x;
    __syncthreads();
    for (int k = 0; k < n; ++k) {
        const int buf = k & 1, wk = k & (LU_TILE_W - 1), ck = k / LU_TILE_W, lk = k & 31, rk = k >> 5;
        if (warp == wk) {  // this warp holds column k
            switch (ck) {
                case 0: LU_COLSTEP(0) break;
                case 1: LU_COLSTEP(1) break;
                case 2: LU_COLSTEP(2) break;
                case 3: LU_COLSTEP(3) break;
                case 4: LU_COLSTEP(4) break;
                case 5: LU_COLSTEP(5) break;
                case 6: LU_COLSTEP(6) break;
                default: LU_COLSTEP(7) break;
            }
        }
        __syncthreads();  // the only barrier of the step: lcol / s_p are double-buffered
        const int p = s_p[buf];
        const int lp = p & 31, rp = p >> 5;
        double l[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int i = lane + 32 * r;
            l[r] = (i > k && i < n) ? lcol[buf][i] : 0.0;
        }
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const int j = warp + LU_TILE_W * c;
            if (j >= n || (warp == wk && c == ck)) continue;  // warp-uniform
            const double rowk = LU_PICKROW(c, rk), rowp = LU_PICKROW(c, rp);
            const double vk = __shfl_sync(0xffffffffu, rowk, lk);
            const double vp = __shfl_sync(0xffffffffu, rowp, lp);  // pivot row element u[k][j]
            if (p != k) {
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    if (r == rk && lane == lk) a[r][c] = vp;
                    if (r == rp && lane == lp) a[r][c] = vk;
                }
            }
            if (j > k) {
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const int i = lane + 32 * r;
                    if (i > k && i < n) a[r][c] = a[r][c] - l[r] * vp;
                }
            }
        }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const int i = lane + 32 * r, j = warp + LU_TILE_W * c;
            if (i < n && j < n) q.LU[(size_t)i * q.lda + j] = a[r][c];
        }
    if (threadIdx.x < n) q.piv[threadIdx.x] = s_piv[threadIdx.x];
}

// Inverse columns: block b owns columns [32 b, 32 b + 32), lane = column.  Shared memory: L\U (n x lda) + x (n x 33).
#define INV_T 128
struct InvColsArgs {
    const double* LU;
    const int* piv;
    int n, lda;
    double* X;  // out: inverse, n x lda row-major
};
__global__ void __launch_bounds__(INV_T, 1) k_inv_cols(InvColsArgs q) {
    extern __shared__ __align__(16) double inv_smem[];
    const int n = q.n, lda = q.lda;
    double* __restrict__ A = inv_smem;
    double* __restrict__ xs = inv_smem + (size_t)n * lda;  // x[i][lane], stride 33
    for (int e = threadIdx.x; e < n * lda; e += INV_T) A[e] = __ldg(q.LU + e);
    __syncthreads();
    if (threadIdx.x >= 32) return;
    const int lane = threadIdx.x, c = blockIdx.x * 32 + lane;
    double* __restrict__ x = xs + lane;
    for (int i = 0; i < n; ++i) x[(size_t)i * 33] = (__ldg(q.piv + i) == c) ? 1.0 : 0.0;  // column c of P * I
    for (int i = 1; i < n; ++i) {  // forward substitution, unit lower triangle, ascending j
        double xi = x[(size_t)i * 33];
        const double* __restrict__ ai = A + (size_t)i * lda;
        int j = 0;
        for (; j + 8 <= i; j += 8) {
            double pr[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) pr[u] = ai[j + u] * x[(size_t)(j + u) * 33];
#pragma unroll
            for (int u = 0; u < 8; ++u) xi = xi - pr[u];
        }
        for (; j < i; ++j) xi = xi - ai[j] * x[(size_t)j * 33];
        x[(size_t)i * 33] = xi;
    }
    for (int i = n - 1; i >= 0; --i) {  // back substitution, ascending j, division last
        double xi = x[(size_t)i * 33];
        const double* __restrict__ ai = A + (size_t)i * lda;
        int j = i + 1;
        for (; j + 8 <= n; j += 8) {
            double pr[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) pr[u] = ai[j + u] * x[(size_t)(j + u) * 33];
#pragma unroll
            for (int u = 0; u < 8; ++u) xi = xi - pr[u];
        }
        for (; j < n; ++j) xi = xi - ai[j] * x[(size_t)j * 33];
        x[(size_t)i * 33] = xi / ai[i];
    }
    if (c < n)
        for (int i = 0; i < n; ++i) q.X[(size_t)i * lda + c] = x[(size_t)i * 33];
}

struct StepFinArgs {
    const double* hg;
    const double* X;
    int n, lda;
    double alpha, max_step;
    double* step;
    double* step2;
    double* tail;
};
__global__ void __launch_bounds__(LU_TILE_N, 1) k_step_fin(StepFinArgs q) {
    __shared__ double s_red[2][LU_TILE_N / 32];
    __shared__ int s_nan;
    const int n = q.n, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const double INF = __longlong_as_double(0x7ff0000000000000ll);
    if (tid == 0) s_nan = 0;
    __syncthreads();
    const double* __restrict__ g = q.hg + (size_t)n * n;
    double s = 0.0, mx = -INF, mn = INF;
    if (tid < n) {
        const double na = -q.alpha;
        const double* __restrict__ xr = q.X + (size_t)tid * q.lda;
        for (int b = 0; b < n; ++b) s += (na * xr[b]) * g[b];
        if (s != s) s_nan = 1;
        mx = s;
        mn = s;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    }
    if (lane == 0) {
        s_red[0][wid] = mx;
        s_red[1][wid] = mn;
    }
    __syncthreads();
    if (tid == 0) q.tail[0] = q.hg[(size_t)n * n + n];
    const bool nan = s_nan != 0;
    if (!nan) {
        for (int w = 0; w < LU_TILE_N / 32; ++w) {
            mx = fmax(mx, s_red[0][w]);
            mn = fmin(mn, s_red[1][w]);
        }
        const double maxElem = fmax(mx, -mn);
        if (maxElem > q.max_step) s = (q.max_step / maxElem) * s;
    }
    if (tid < n) {
        q.step[tid] = s;
        if (q.step2) q.step2[tid] = s;
    }
    if (tid == 0) q.tail[1] = nan ? 1.0 : 0.0;
}

}  // namespace dmsa
