// kernels_solve.cuh — the Levenberg-Marquardt step of DmsaOptimizer.h:107-128 on the device (P <= 128), so that an
// iteration does not stop for a host round trip between the Jacobian pass and the line search:
//
//   H.diagonal().array() += lambda_diag                                   :108-111
//   step = (-alpha * H.inverse()) * (J^T e0)                              :113  (Eigen dynamic inverse() == PartialPivLU)
//   NaN guard                                                             :115-122
//   if (max|step| > max_step) step = (max_step / max|step|) * step        :124-128
//
// The arithmetic is the operation sequence of host_solve.cpp (lu_factor_impl / lu_subst_block_impl / dmsa_host_lm_step),
// element for element: LU with partial pivoting (first maximum wins), explicit inverse by substituting the n columns of
// P * I — forward with ascending j, backward column-oriented with descending j and a multiplication by the reciprocal
// diagonal — then the matrix-vector product accumulated in ascending column order.  IEEE double, no FMA contraction
// (the library is built with -fmad=false), so host and device produce bit-identical steps
// (tests/test_gpu_parity.py::test_device_lm_solve_is_bit_identical_to_the_host_solver).
//
// Three kernels:
//   k_lu128     one block, the matrix in registers (thread (lane, warp) holds rows lane + 32 r, columns 8 warp + c), rows stay
//               in place (a position table replaces the physical row swaps), panels of 8 columns, ONE barrier per panel;
//   k_inv128    the n right-hand sides are independent: one warp per column of the inverse, L\U staged column-major in
//               shared memory, per step one shuffle broadcast + a multiply-subtract per row (chain of 2 n steps);
//   k_step_fin  step = (-alpha X) g, NaN flag, infinity-norm clamp (products in parallel, the additions of a row in order).
#pragma once
#include <cuda_runtime.h>

#include "pdl.cuh"

namespace dmsa {

#define LM_DEV_MAXN 128  // largest system the device solver takes (larger ones: host solver)
#define LU128_T 512      // 16 warps: warp w holds the 8 columns 8 w .. 8 w + 7, lane l the rows l + 32 r (r < 4)
#define LU128_PANEL 8

__device__ long long g_lu_clk[16];
#define LUCLK(i) if (DBG && pnl == 1 && kk == 3 && lane == 0) g_lu_clk[i] = clock64();
struct Lu128Args {
    const double* hg;  // [H (n x n row-major) | g | err0]
    int n, ld;         // ld: leading dimension of the column-major output (multiple of 32, >= n)
    double lambda;
    double* LUT;       // out: L\U TRANSPOSED (column-major): LUT[j * ld + i] = LU[i][j], rows in pivoted order
    double* rdiag;     // out: 1 / U[j][j]
    int* piv;          // out: piv[i] = original row now at position i
};

// a[RP][C] of lane LP for every lane: RP is warp-uniform, so the register row is picked by a uniform branch (no selects)
#define LU_ROWBCAST(DST, C, RP, LP)                                  \
    switch (RP) {                                                    \
        case 0: DST = __shfl_sync(0xffffffffu, a[0][C], LP); break;  \
        case 1: DST = __shfl_sync(0xffffffffu, a[1][C], LP); break;  \
        case 2: DST = __shfl_sync(0xffffffffu, a[2][C], LP); break;  \
        default: DST = __shfl_sync(0xffffffffu, a[3][C], LP); break; \
    }

// Correctly rounded x / d for the four rows of a lane with ONE reciprocal per step (Markstein: y = RN(1 / d),
// q = RN(x y), r = x - d q exactly (FMA), q' = RN(q + r y) == RN(x / d) when nothing over- or underflows and the
// significand of d is not all ones); `safe` (warp-uniform) says that every operand of the warp is inside that envelope,
// otherwise the plain IEEE division runs.  Either way the result is the IEEE quotient the host computes.
__device__ __forceinline__ bool lu_div_operand_safe(double x) {  // zero, or a normal number whose quotients stay far from
    const unsigned e = ((unsigned)(__double_as_longlong(x) >> 52)) & 0x7ffu;  // the exponent limits (2^-423 < |x| < 2^377)
    return x == 0.0 || (e > 600u && e < 1400u);
}

// Right-looking LU with partial pivoting, the matrix in registers.  Rows never move: a position table replaces the row
// exchanges (the arithmetic of an element does not depend on where its row is stored).  A panel = the 8 columns of one warp:
// its owner factorises it inside the warp (pivot search = three warp reductions on the bit pattern of |a|, first maximum
// in position order; pivot row broadcast by shuffles), publishes the multipliers and pivots of its 8 steps in shared
// memory, and after ONE block barrier per panel every warp with columns to the right applies the 8 elimination steps to
// its registers (each element sees the steps in ascending order, a multiply and a subtract per step: the sequence of an
// unblocked elimination).
template <bool DBG>
__global__ void __launch_bounds__(LU128_T, 1) k_lu128(Lu128Args q) {
    DMSA_PDL_ENTER();
    __shared__ double f_s[2][LU128_PANEL][LM_DEV_MAXN];  // multipliers of the panel's steps, by physical row
    __shared__ int pinfo[2][LU128_PANEL];                // pivot of each step: (position it came from) * 128 + physical row
    __shared__ int s_piv[LM_DEV_MAXN];
    __shared__ int s_fpos[LM_DEV_MAXN];
    const int n = q.n, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double a[4][8];  // a[r][c] = A[lane + 32 r][8 warp + c] (physical rows)
    int pos[4];      // current position of physical row lane + 32 r in the pivoted order
    unsigned done = 0;  // bit r: row is a finished pivot row (or does not exist)
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int i = lane + 32 * r;
        pos[r] = i;
        if (i >= n) done |= 1u << r;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const int j = 8 * warp + c;
            double v = 0.0;
            if (i < n && j < n) {
                v = __ldg(q.hg + (size_t)i * n + j);
                if (i == j) v += q.lambda;
            }
            a[r][c] = v;
        }
    }
    const int npanel = (n + LU128_PANEL - 1) / LU128_PANEL;
    for (int pnl = 0; pnl < npanel; ++pnl) {
        const int buf = pnl & 1;
        if (warp == pnl) {
            // ---- this warp holds the panel: factorise its 8 columns ----
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) {
                const int k = 8 * pnl + kk;
                if (k < n) {  // warp-uniform
                    LUCLK(0)
                    // pivot: first maximum of |a| in position order.  Non-negative doubles order like their bit patterns:
                    // maximum of the high words, then of the low words among the lanes that hold it, then the smallest
                    // (position, row) key among those.  (A NaN in the column orders above everything here, while the host's
                    // `v > best` never lets one replace the running best: either way a NaN in the matrix ends as a NaN
                    // step and the optimizer's NaN guard - the flag is what host and device agree on, not the NaN pattern.)
                    unsigned long long bb = 0ull;
                    int bkey = 0x7fffffff;
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        if (!((done >> r) & 1u)) {
                            const unsigned long long vb = (unsigned long long)__double_as_longlong(a[r][kk]) & 0x7fffffffffffffffull;
                            const int key = pos[r] * 128 + lane + 32 * r;
                            if (vb > bb || (vb == bb && key < bkey)) {
                                bb = vb;
                                bkey = key;
                            }
                        }
                    }
                    const unsigned hi = (unsigned)(bb >> 32), lo = (unsigned)bb;
                    const unsigned mhi = __reduce_max_sync(0xffffffffu, hi);
                    const unsigned mlo = __reduce_max_sync(0xffffffffu, hi == mhi ? lo : 0u);
                    bkey = __reduce_min_sync(0xffffffffu, (hi == mhi && lo == mlo) ? bkey : 0x7fffffff);
                    const int p = bkey & 127, ppos = bkey >> 7, rp = p >> 5, lp = p & 31;
                    LUCLK(1)
                    double d;
                    LU_ROWBCAST(d, kk, rp, lp)
                    LUCLK(2)
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        const int i = lane + 32 * r;
                        if (i == p) {
                            pos[r] = k;
                            done |= 1u << r;
                        } else if (pos[r] == k) {
                            pos[r] = ppos;  // the row that sat at position k moves to where the pivot row came from
                        }
                    }
                    // multipliers f = a / d of the unfinished rows
                    bool safe = lu_div_operand_safe(d) && d != 0.0 && (__double_as_longlong(d) & 0xfffffffffffffll) != 0xfffffffffffffll;
#pragma unroll
                    for (int r = 0; r < 4; ++r)
                        if (!((done >> r) & 1u)) safe = safe && lu_div_operand_safe(a[r][kk]);
                    safe = __all_sync(0xffffffffu, safe);
                    double f[4];
                    if (safe) {
                        const double y = 1.0 / d;
#pragma unroll
                        for (int r = 0; r < 4; ++r) {
                            const double x = a[r][kk];
                            const double q0 = __dmul_rn(x, y);
                            const double rem = __fma_rn(-d, q0, x);
                            f[r] = (x == 0.0) ? q0 : __fma_rn(rem, y, q0);  // (a zero keeps the sign of the IEEE quotient)
                        }
                    } else {
#pragma unroll
                        for (int r = 0; r < 4; ++r) f[r] = a[r][kk] / d;
                    }
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        if (!((done >> r) & 1u)) {
                            a[r][kk] = f[r];
                            f_s[buf][kk][lane + 32 * r] = f[r];
                        }
                    }
                    LUCLK(3)
                    if (kk < 7) {
                        double u[8];
                        switch (rp) {  // warp-uniform
                            case 0:
#pragma unroll
                                for (int c = kk + 1; c < 8; ++c) u[c] = __shfl_sync(0xffffffffu, a[0][c], lp);
                                break;
                            case 1:
#pragma unroll
                                for (int c = kk + 1; c < 8; ++c) u[c] = __shfl_sync(0xffffffffu, a[1][c], lp);
                                break;
                            case 2:
#pragma unroll
                                for (int c = kk + 1; c < 8; ++c) u[c] = __shfl_sync(0xffffffffu, a[2][c], lp);
                                break;
                            default:
#pragma unroll
                                for (int c = kk + 1; c < 8; ++c) u[c] = __shfl_sync(0xffffffffu, a[3][c], lp);
                                break;
                        }
#pragma unroll
                        for (int r = 0; r < 4; ++r)
                            if (!((done >> r) & 1u)) {
#pragma unroll
                                for (int c = kk + 1; c < 8; ++c) a[r][c] = __dsub_rn(a[r][c], __dmul_rn(f[r], u[c]));
                            }
                    }
                    if (lane == 0) {
                        pinfo[buf][kk] = bkey;
                        s_piv[k] = p;
                    }
                    LUCLK(4)
                }
            }
        }
        if (DBG && pnl == 1 && threadIdx.x == 32 * 2) g_lu_clk[5] = clock64();  // warp 2 (the next owner) reaches the barrier
        __syncthreads();  // the only block barrier of the panel (f_s / pinfo are double-buffered)
        if (warp > pnl && 8 * warp < n) {
            // ---- trailing update of this warp's columns with the panel's steps ----
#pragma unroll 1
            for (int kk = 0; kk < 8; ++kk) {
                const int k = 8 * pnl + kk;
                if (k >= n) break;
                const int key = pinfo[buf][kk];
                if (DBG && pnl == 1 && threadIdx.x == 32 * 2) g_lu_clk[8 + kk] = clock64();
                const int p = key & 127, ppos = key >> 7, rp = p >> 5, lp = p & 31;
                double f[4];
#pragma unroll
                for (int r = 0; r < 4; ++r) f[r] = f_s[buf][kk][lane + 32 * r];
                double u[8];
                switch (rp) {  // warp-uniform
                    case 0:
#pragma unroll
                        for (int c = 0; c < 8; ++c) u[c] = __shfl_sync(0xffffffffu, a[0][c], lp);
                        break;
                    case 1:
#pragma unroll
                        for (int c = 0; c < 8; ++c) u[c] = __shfl_sync(0xffffffffu, a[1][c], lp);
                        break;
                    case 2:
#pragma unroll
                        for (int c = 0; c < 8; ++c) u[c] = __shfl_sync(0xffffffffu, a[2][c], lp);
                        break;
                    default:
#pragma unroll
                        for (int c = 0; c < 8; ++c) u[c] = __shfl_sync(0xffffffffu, a[3][c], lp);
                        break;
                }
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const int i = lane + 32 * r;
                    if (i == p) {
                        pos[r] = k;
                        done |= 1u << r;
                    } else if (pos[r] == k) {
                        pos[r] = ppos;
                    }
                }
#pragma unroll
                for (int r = 0; r < 4; ++r)
                    if (!((done >> r) & 1u)) {
                        double t[8];
#pragma unroll
                        for (int c = 0; c < 8; ++c) t[c] = __dmul_rn(f[r], u[c]);
#pragma unroll
                        for (int c = 0; c < 8; ++c) a[r][c] = __dsub_rn(a[r][c], t[c]);
                    }
            }
            if (DBG && pnl == 1 && threadIdx.x == 32 * 2) g_lu_clk[7] = clock64();  // warp 2 finished its trailing update of panel 1
        }
    }
    __syncthreads();
    // rows leave in pivoted order (final position of physical row i = the step that made it the pivot row), transposed
    // (column-major) for the substitution kernel
    if (threadIdx.x < n) s_fpos[s_piv[threadIdx.x]] = threadIdx.x;
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int i = lane + 32 * r;
        if (i >= n) continue;
        const int ip = s_fpos[i];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const int j = 8 * warp + c;
            if (j < n) {
                q.LUT[(size_t)j * q.ld + ip] = a[r][c];
                if (ip == j) q.rdiag[j] = 1.0 / a[r][c];
            }
        }
    }
    if (threadIdx.x < n) q.piv[threadIdx.x] = s_piv[threadIdx.x];
}

// Columns of the inverse: one warp per right-hand side (column c of P * I), lane l holds rows l + 32 r.
#define INV128_WARPS 4
struct Inv128Args {
    const double* LUT;    // column-major L\U
    const double* rdiag;  // 1 / U[j][j]
    const int* piv;
    int n, ld;
    double* XT;           // out: the inverse TRANSPOSED: XT[c * ld + i] = X[i][c]
};
__global__ void __launch_bounds__(INV128_WARPS * 32, 1) k_inv128(Inv128Args q) {
    DMSA_PDL_ENTER();
    extern __shared__ __align__(16) double inv_s[];  // LUT (n x ld) | rdiag (n)
    const int n = q.n, ld = q.ld, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* __restrict__ S = inv_s;
    double* __restrict__ rd = inv_s + (size_t)n * ld;
    {
        const double2* __restrict__ src = reinterpret_cast<const double2*>(q.LUT);
        double2* __restrict__ dst = reinterpret_cast<double2*>(S);
        const int tot2 = n * ld / 2;
        for (int e = threadIdx.x; e < tot2; e += INV128_WARPS * 32) dst[e] = __ldg(src + e);
        for (int e = threadIdx.x; e < n; e += INV128_WARPS * 32) rd[e] = __ldg(q.rdiag + e);
    }
    __syncthreads();
    const int col = blockIdx.x * INV128_WARPS + warp;
    if (col >= n) return;
    double x[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int i = lane + 32 * r;
        x[r] = (i < n && __ldg(q.piv + i) == col) ? 1.0 : 0.0;  // column `col` of P * I
    }
    // forward substitution, unit lower triangle: x_i -= l_ij x_j for ascending j
#pragma unroll
    for (int rj = 0; rj < 4; ++rj) {
        for (int jj = 0; jj < 32; ++jj) {
            const int j = 32 * rj + jj;
            if (j >= n) break;
            const double xj = __shfl_sync(0xffffffffu, x[rj], jj);
            const double* __restrict__ Lc = S + (size_t)j * ld + lane;
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                if (r < rj) continue;  // compile-time
                const int i = lane + 32 * r;
                if (i > j && i < n) x[r] = __dsub_rn(x[r], __dmul_rn(Lc[32 * r], xj));
            }
        }
    }
    // backward substitution, column-oriented: x_j *= 1 / u_jj, then x_i -= u_ij x_j for the rows above, descending j
#pragma unroll
    for (int rj = 3; rj >= 0; --rj) {
        for (int jj = 31; jj >= 0; --jj) {
            const int j = 32 * rj + jj;
            if (j >= n) continue;
            if (lane == jj) x[rj] = __dmul_rn(x[rj], rd[j]);
            const double xj = __shfl_sync(0xffffffffu, x[rj], jj);
            const double* __restrict__ Uc = S + (size_t)j * ld + lane;
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                if (r > rj) continue;  // compile-time
                const int i = lane + 32 * r;
                if (i < j) x[r] = __dsub_rn(x[r], __dmul_rn(Uc[32 * r], xj));
            }
        }
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int i = lane + 32 * r;
        if (i < n) q.XT[(size_t)col * ld + i] = x[r];
    }
}

struct StepFinArgs {
    const double* hg;
    const double* XT;  // transposed inverse (k_inv128)
    int n, ld;
    double alpha, max_step;
    double* step;
    double* step2;     // second copy inside the iteration's read-back block (may be null)
    double* tail;      // out [0] = err0 (copied from hg), [1] = 1.0 if the step contains NaN (then left unclamped), else 0.0
};
// The n^2 products (-alpha X[i][b]) g[b] are independent: 512 threads form them with coalesced loads (one L2 round trip)
// and leave them in shared memory; only the n additions of a row are a dependent chain (ascending b, like the host).
#define STEPFIN_T 512
__global__ void __launch_bounds__(STEPFIN_T, 1) k_step_fin(StepFinArgs q) {
    DMSA_PDL_ENTER();
    extern __shared__ __align__(16) double sf_t[];  // [b][ld] products
    __shared__ double s_red[2][LM_DEV_MAXN / 32];
    __shared__ double s_g[LM_DEV_MAXN];
    __shared__ int s_nan;
    const int n = q.n, ld = q.ld, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const double INF = __longlong_as_double(0x7ff0000000000000ll);
    if (tid == 0) s_nan = 0;
    if (tid < n) s_g[tid] = q.hg[(size_t)n * n + tid];
    __syncthreads();
    {
        const double na = -q.alpha;
        const int i = tid & (LM_DEV_MAXN - 1), b0 = tid / LM_DEV_MAXN;  // X[i][b] = XT[b * ld + i]: coalesced over the threads
        if (i < n) {
#pragma unroll 8
            for (int b = b0; b < n; b += STEPFIN_T / LM_DEV_MAXN) sf_t[(size_t)b * ld + i] = __dmul_rn(__dmul_rn(na, q.XT[(size_t)b * ld + i]), s_g[b]);
        }
    }
    __syncthreads();
    double s = 0.0, mx = -INF, mn = INF;
    if (tid < n) {
        const double* __restrict__ tr = sf_t + tid;
#pragma unroll 8
        for (int b = 0; b < n; ++b) s = __dadd_rn(s, tr[(size_t)b * ld]);
        if (s != s) s_nan = 1;
        mx = s;
        mn = s;
    }
    if (tid < LM_DEV_MAXN) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        }
        if (lane == 0) {
            s_red[0][wid] = mx;
            s_red[1][wid] = mn;
        }
    }
    __syncthreads();
    if (tid == 0) q.tail[0] = q.hg[(size_t)n * n + n];
    const bool nan = s_nan != 0;
    if (!nan) {
        for (int w = 0; w < LM_DEV_MAXN / 32; ++w) {
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
