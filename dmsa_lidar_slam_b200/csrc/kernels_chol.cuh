// kernels_chol.cuh — LM step for the keyframe-BUNDLE extension (BASELINE config 4): the all-reduced global system
// (H + lambda I) x = g of the bundles' scattered J^T J blocks (P = 378 for 64 keyframes) is symmetric positive definite, and
// the bundle scheme has no reference arithmetic to mirror (the reference optimises ONE submap, DmsaSlam.h:212-238), so it is
// solved by a blocked Cholesky factorisation instead of the explicit LU inverse of DmsaOptimizer.h:113:
//     step = -alpha x,  NaN guard,  infinity-norm clamp                                   DmsaOptimizer.h:113-128
// One cooperative kernel (grid-wide barriers between the phases of a block column): 32 x 32 diagonal block factorised by
// one warp with shuffles, the block column below it solved row by row (one thread per row), the trailing matrix updated
// tile by tile over the grid.  The right-hand side rides along as an extra matrix row (row n of the factor is
// y = L^-1 g), so only the back substitution L^T x = y remains, done block-wise by one block.  Deterministic (fixed
// operation order, no atomics): every rank computes the same step from the same all-reduced system.
#pragma once
#include <cooperative_groups.h>
#include <cuda_runtime.h>

#include "kernels_sort.cuh"  // DMSA_TLK (development time stamps)

namespace dmsa {

#define CHOL_B 32
#define CHOL_T 256
#define CHOL_MAXN 1024

struct CholArgs {
    const double* hg;  // [H (n x n row-major) | g (n) | err0]
    int n, ld;         // ld >= n, multiple of 32
    double lambda, alpha, max_step;
    double* W;         // work: (n + 1) x ld, lower triangle of H + lambda I, row n = g
    double* step;      // out [n]
    double* step2;     // out [n] second copy (may be null)
    double* tail;      // out [0] = err0, [1] = 0 ok / 1 NaN step / 2 not positive definite (caller falls back)
    int* flag;         // device scratch (1 int, zeroed by the kernel)
    double* dinv;      // device scratch [ld]: reciprocals of the factor's diagonal
};

// 32 x 32 diagonal block at W[j0.., j0..] -> its Cholesky factor in place (one warp; lane = row).  nb = valid rows (<= 32).
// dinv[c] = 1 / L[c][c] for the substitutions that follow.  The square root and the division of the textbook step are ONE
// reciprocal square root (l = d * rsqrt(d)): the two IEEE sequences were 90 % of this function's 30 us (measured with the
// phase stamps of scripts/dbg_chol_timeline.py); the bundle scheme has no reference arithmetic to mirror, only determinism.
__device__ __noinline__ void chol_diag_block(double* __restrict__ Wjj, int ld, int nb, int lane, int* flag, double* __restrict__ dinv,
                                             double* __restrict__ sL /* shared, CHOL_B doubles */) {
    double a[CHOL_B];
#pragma unroll
    for (int c = 0; c < CHOL_B; ++c) a[c] = (lane < nb && c <= lane) ? Wjj[(size_t)lane * ld + c] : (c == lane ? 1.0 : 0.0);
    bool bad = false;
    double myinv = 1.0;
#pragma unroll
    for (int c = 0; c < CHOL_B; ++c) {
        const double d = __shfl_sync(0xffffffffu, a[c], c);
        if (!(d > 0.0)) bad = true;
        const double inv = rsqrt(d);
        if (lane == c) {
            a[c] = d * inv;
            myinv = inv;
        }
        if (lane > c) a[c] = a[c] * inv;
        sL[lane] = a[c];  // column c of the factor (rows above the diagonal hold zeros / the unit padding)
        __syncwarp();
#pragma unroll
        for (int c2 = c + 1; c2 < CHOL_B; ++c2)
            if (lane >= c2) a[c2] = fma(-a[c], sL[c2], a[c2]);
        __syncwarp();
    }
    if (bad && lane == 0) *flag = 2;
    if (lane < nb) dinv[lane] = myinv;
#pragma unroll
    for (int c = 0; c < CHOL_B; ++c)
        if (lane < nb && c <= lane) Wjj[(size_t)lane * ld + c] = a[c];
}

// one row of the block column below a diagonal block, by ONE WARP: x L_jj^T = w with lane k owning x[k]; per step the
// finished x[c] is broadcast and every later lane takes its term (32 short steps; one thread per row walked 496 dependent
// multiply-adds and took 14 us per block column)
__device__ __forceinline__ void chol_panel_row_warp(double* __restrict__ wr, int nbj, const double (*sA)[CHOL_B + 1], const double* sInv, int lane) {
    double x = lane < nbj ? wr[lane] : 0.0;
#pragma unroll
    for (int c = 0; c < CHOL_B; ++c) {
        const double xc = __shfl_sync(0xffffffffu, x, c) * sInv[c];
        if (lane == c) x = xc;
        if (lane > c) x = fma(-xc, sA[lane][c], x);
    }
    if (lane < nbj) wr[lane] = x;
}

__global__ void __launch_bounds__(CHOL_T, 1) k_chol_solve(CholArgs q) {
    namespace cg = cooperative_groups;
    cg::grid_group grid = cg::this_grid();
    __shared__ double sA[CHOL_B][CHOL_B + 1], sB[CHOL_B][CHOL_B + 1];
    __shared__ double sx[CHOL_MAXN];
    __shared__ double s_red[2][CHOL_T / 32];
    __shared__ int s_nan;
    const int n = q.n, ld = q.ld, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int gthreads = gridDim.x * CHOL_T, gtid = blockIdx.x * CHOL_T + tid;
    double* __restrict__ W = q.W;
    // load: lower triangle of H + lambda I, row n = g
    for (int e = gtid; e < (n + 1) * n; e += gthreads) {
        const int i = e / n, j = e - i * n;
        double v = 0.0;
        if (i < n) {
            if (j <= i) v = q.hg[(size_t)i * n + j] + (i == j ? q.lambda : 0.0);
        } else {
            v = q.hg[(size_t)n * n + j];
        }
        W[(size_t)i * ld + j] = v;
    }
    DMSA_TLK(3, 0);
    if (gtid == 0) *q.flag = 0;
    grid.sync();
    DMSA_TLK(3, 1);
    const int nblk = (n + CHOL_B - 1) / CHOL_B;
    const int nrows = n + 1;  // the right-hand side is row n
    for (int jb = 0; jb < nblk; ++jb) {
        const int j0 = jb * CHOL_B, nbj = min(CHOL_B, n - j0);
        // (1) diagonal block
        if (blockIdx.x == 0 && warp == 0) chol_diag_block(W + (size_t)j0 * ld + j0, ld, nbj, lane, q.flag, q.dinv + j0, &sB[0][0]);
        if (jb == 0) DMSA_TLK(3, 2);
        grid.sync();
        if (jb == 0) DMSA_TLK(3, 3);
        // (2) rows below (and the right-hand-side row): x L_jj^T = w, one thread per row, right-looking (as soon as x[c] is known
        // every later entry takes its term: 32 short dependent steps instead of a chain of 496 multiply-adds and 32 divisions)
        {
            __shared__ double sInv[CHOL_B];
            for (int e = tid; e < CHOL_B * CHOL_B; e += CHOL_T) {
                const int r = e / CHOL_B, c = e - r * CHOL_B;
                sA[r][c] = (r < nbj && c <= r) ? W[(size_t)(j0 + r) * ld + j0 + c] : 0.0;
            }
            if (tid < CHOL_B) sInv[tid] = tid < nbj ? q.dinv[j0 + tid] : 1.0;
            __syncthreads();
            const int r0 = j0 + nbj;  // first row below the diagonal block
            const int gwarp = blockIdx.x * (CHOL_T / 32) + warp, gwarps = gridDim.x * (CHOL_T / 32);
            for (int r = r0 + gwarp; r < nrows; r += gwarps) chol_panel_row_warp(W + (size_t)r * ld + j0, nbj, sA, sInv, lane);
        }
        if (jb == 0) DMSA_TLK(3, 4);
        grid.sync();
        // (3) trailing update: W[bi][bk] -= W[bi][jb] W[bk][jb]^T for jb < bk <= bi (row blocks run to row n)
        {
            const int first = jb + 1;
            const int nrb = (nrows - first * CHOL_B + CHOL_B - 1) / CHOL_B;  // row blocks below
            const int ncb = nblk - first;                                   // column blocks to the right
            // tiles (bi, bk), bk in [0, ncb), bi in [bk, nrb): enumerate bi-major
            int ntile = 0;
            for (int bk = 0; bk < ncb; ++bk) ntile += nrb - bk;
            for (int t = blockIdx.x; t < ntile; t += gridDim.x) {
                int bk = 0, rem = t;
                while (rem >= nrb - bk) {
                    rem -= nrb - bk;
                    ++bk;
                }
                const int bi = bk + rem;
                const int i0 = (first + bi) * CHOL_B, k0 = (first + bk) * CHOL_B;
                __syncthreads();
                for (int e = tid; e < CHOL_B * CHOL_B; e += CHOL_T) {
                    const int r = e / CHOL_B, c = e - r * CHOL_B;
                    sA[r][c] = (i0 + r < nrows && c < nbj) ? W[(size_t)(i0 + r) * ld + j0 + c] : 0.0;
                    sB[r][c] = (k0 + r < n && c < nbj) ? W[(size_t)(k0 + r) * ld + j0 + c] : 0.0;
                }
                __syncthreads();
                // 256 threads: thread -> output (r, 4 columns)
                const int r = tid >> 3, cq = (tid & 7) * 4;
                double acc[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll 8
                for (int k = 0; k < CHOL_B; ++k) {
                    const double av = sA[r][k];
#pragma unroll
                    for (int u = 0; u < 4; ++u) acc[u] = fma(av, sB[cq + u][k], acc[u]);
                }
                if (i0 + r < nrows) {
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int col = k0 + cq + u;
                        if (col < n && col <= i0 + r) W[(size_t)(i0 + r) * ld + col] -= acc[u];
                    }
                }
            }
        }
        if (jb == 0) DMSA_TLK(3, 5);
        grid.sync();
    }
    DMSA_TLK(3, 6);
    if (blockIdx.x != 0) return;
    // back substitution L^T x = y (y = row n), block-wise from the bottom: one block
    for (int i = tid; i < n; i += CHOL_T) sx[i] = W[(size_t)n * ld + i];
    if (tid == 0) s_nan = 0;
    __syncthreads();
    for (int jb = nblk - 1; jb >= 0; --jb) {
        const int j0 = jb * CHOL_B, nbj = min(CHOL_B, n - j0);
        for (int e = tid; e < CHOL_B * CHOL_B; e += CHOL_T) {  // the diagonal block, coalesced
            const int r = e / CHOL_B, c = e - r * CHOL_B;
            sA[r][c] = (r < nbj && c <= r) ? W[(size_t)(j0 + r) * ld + j0 + c] : 0.0;
        }
        __syncthreads();
        if (warp == 0) {  // 32 x 32 triangular solve L_jj^T x_j = rhs_j: column-oriented, descending
            double xv = lane < nbj ? sx[j0 + lane] : 0.0;
            const double myinv = lane < nbj ? q.dinv[j0 + lane] : 1.0;
            for (int c = nbj - 1; c >= 0; --c) {
                const double xc = __shfl_sync(0xffffffffu, xv, c) * __shfl_sync(0xffffffffu, myinv, c);
                if (lane == c) xv = xc;
                if (lane < c) xv = fma(-sA[c][lane], xc, xv);
            }
            if (lane < nbj) sx[j0 + lane] = xv;
        }
        __syncthreads();
        // rhs_k -= sum_c L[j0 + c][k] x[j0 + c] for k < j0 (32 independent coalesced loads per thread)
        for (int k = tid; k < j0; k += CHOL_T) {
            double lv[CHOL_B];
#pragma unroll
            for (int c = 0; c < CHOL_B; ++c) lv[c] = c < nbj ? W[(size_t)(j0 + c) * ld + k] : 0.0;
            double s = sx[k];
#pragma unroll
            for (int c = 0; c < CHOL_B; ++c) s = fma(-lv[c], sx[j0 + c < n ? j0 + c : j0], s);
            sx[k] = s;
        }
        __syncthreads();
    }
    // step = -alpha x, NaN guard, infinity-norm clamp (DmsaOptimizer.h:113-128)
    const double INF = __longlong_as_double(0x7ff0000000000000ll);
    double mx = -INF, mn = INF;
    for (int i = tid; i < n; i += CHOL_T) {
        const double s = -q.alpha * sx[i];
        sx[i] = s;
        if (s != s) s_nan = 1;
        mx = fmax(mx, s);
        mn = fmin(mn, s);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    }
    if (lane == 0) {
        s_red[0][warp] = mx;
        s_red[1][warp] = mn;
    }
    __syncthreads();
    for (int w = 0; w < CHOL_T / 32; ++w) {
        mx = fmax(mx, s_red[0][w]);
        mn = fmin(mn, s_red[1][w]);
    }
    const bool nan = s_nan != 0;
    const double maxElem = fmax(mx, -mn);
    const double scale = (!nan && maxElem > q.max_step) ? q.max_step / maxElem : 1.0;
    for (int i = tid; i < n; i += CHOL_T) {
        const double s = scale == 1.0 ? sx[i] : scale * sx[i];
        q.step[i] = s;
        if (q.step2) q.step2[i] = s;
    }
    DMSA_TLK(3, 7);
    if (tid == 0) {
        q.tail[0] = q.hg[(size_t)n * n + n];
        q.tail[1] = (*q.flag == 2) ? 2.0 : (nan ? 1.0 : 0.0);
    }
}

}  // namespace dmsa
