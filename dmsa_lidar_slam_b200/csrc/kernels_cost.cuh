// kernels_cost.cuh — the hot loop: batched cost evaluation for V parameter vectors at once
// (DmsaOptimizer.h:234-273 updateErrorTerms, fused with the per-point transform of
// ContinuousTrajectory.h:137-155 / MapManagement.h:140-147) and the J^T J / J^T e reduction (DmsaOptimizer.h:107,113).
//
// Mapping: one thread block per chunk of <= CH members of one Gaussian set, one thread per parameter vector v.
// Every member record (16 B: local xyz + transform-row index) is read once per pass for ALL V vectors (uniform,
// broadcast load); the 48-byte float transform of (row, v) is fetched only when the row index changes between
// consecutive members.  Float arithmetic is the reference's, operation for operation, with explicit
// round-to-nearest intrinsics (no FMA contraction: the reference is built without FMA):
//   world  = ((m0*x + m1*y) + m2*z) + m3            Matrix4f * Vector4f, w == 1 (checked at upload)
//   mean   = float(sum_j world_j) / float(n)        exactly-rounded sum (see DESIGN.md "mean")
//   d      = world - mean
//   term   = ((w*d)^T * info) * d                   row-vector * Matrix3f * vector, 3-element redux a0 + (a1 + a2)
//   e      = sqrt(|sum_j double(term_j)|)
#pragma once
#include <cuda_runtime.h>

#include "kernels_sets.cuh"

namespace dmsa {

struct CostArgs {
    const Chunk* chunks;
    const int* n_chunks;     // device scalar: total chunks (grid is an upper bound)
    const float4* rec;       // member records, sorted order
    const float4* Mtab;      // [(row * Vld + v) * 3 + r]
    int V, Vld;
    const float* info;       // [g][9]
    const float* w;          // [g]
    const int* cell_n;       // [g]
    const int* nchunk;       // [g]
    const int* chunk_off;    // [g]
    double* S;               // partial sums   [(c*3 + a) * Vld + v]
    float* mu;               // means          [(g*3 + a) * Vld + v]
    double* Q;               // partial quadratic forms [c * Vld + v]
    double* E;               // residuals      [g * Vld + v]
};

__device__ __forceinline__ void load_tform(const float4* __restrict__ Mtab, int row, int Vld, int v, float4& m0, float4& m1, float4& m2) {
    const float4* M = Mtab + ((size_t)row * Vld + v) * 3;
    m0 = __ldg(M);
    m1 = __ldg(M + 1);
    m2 = __ldg(M + 2);
}
__device__ __forceinline__ void xform(const float4& m0, const float4& m1, const float4& m2, const float4& r, float& X, float& Y, float& Z) {
    X = fadd_(fadd_(fadd_(fmul_(m0.x, r.x), fmul_(m0.y, r.y)), fmul_(m0.z, r.z)), m0.w);
    Y = fadd_(fadd_(fadd_(fmul_(m1.x, r.x), fmul_(m1.y, r.y)), fmul_(m1.z, r.z)), m1.w);
    Z = fadd_(fadd_(fadd_(fmul_(m2.x, r.x), fmul_(m2.y, r.y)), fmul_(m2.z, r.z)), m2.w);
}

// pass 1: per-chunk coordinate sums (double)
__global__ void __launch_bounds__(1024) k_cost_sum(CostArgs a) {
    const int c = blockIdx.x;
    if (c >= *a.n_chunks) return;
    const int v = threadIdx.x;
    if (v >= a.V) return;
    const Chunk ch = a.chunks[c];
    const float4* __restrict__ rec = a.rec + ch.start;
    double sx = 0.0, sy = 0.0, sz = 0.0;
    int tprev = -2;
    float4 m0, m1, m2;
    m0 = m1 = m2 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
    for (int j = 0; j < ch.count; ++j) {
        const float4 r = __ldg(rec + j);
        const int t = __float_as_int(r.w);
        float X = r.x, Y = r.y, Z = r.z;
        if (t >= 0) {
            if (t != tprev) {
                load_tform(a.Mtab, t, a.Vld, v, m0, m1, m2);
                tprev = t;
            }
            xform(m0, m1, m2, r, X, Y, Z);
        }
        sx += (double)X;
        sy += (double)Y;
        sz += (double)Z;
    }
    a.S[((size_t)c * 3 + 0) * a.Vld + v] = sx;
    a.S[((size_t)c * 3 + 1) * a.Vld + v] = sy;
    a.S[((size_t)c * 3 + 2) * a.Vld + v] = sz;
}

// per (set, v): mean = float(sum over the set's chunks, in chunk order) / float(n)        DmsaOptimizer.h:254
__global__ void k_cost_mean(CostArgs a, int G) {
    const int g = blockIdx.x;
    const int v = threadIdx.x;
    if (g >= G || v >= a.V) return;
    const int nc = a.nchunk[g];
    if (nc == 0) return;
    const int o = a.chunk_off[g];
    double sx = 0.0, sy = 0.0, sz = 0.0;
    for (int c = 0; c < nc; ++c) {
        sx += a.S[((size_t)(o + c) * 3 + 0) * a.Vld + v];
        sy += a.S[((size_t)(o + c) * 3 + 1) * a.Vld + v];
        sz += a.S[((size_t)(o + c) * 3 + 2) * a.Vld + v];
    }
    const float nf = (float)a.cell_n[g];
    a.mu[((size_t)g * 3 + 0) * a.Vld + v] = fdiv_((float)sx, nf);
    a.mu[((size_t)g * 3 + 1) * a.Vld + v] = fdiv_((float)sy, nf);
    a.mu[((size_t)g * 3 + 2) * a.Vld + v] = fdiv_((float)sz, nf);
}

// pass 2: per-chunk sums of the Mahalanobis terms                                        DmsaOptimizer.h:259-264
__global__ void __launch_bounds__(1024) k_cost_quad(CostArgs a) {
    const int c = blockIdx.x;
    if (c >= *a.n_chunks) return;
    const int v = threadIdx.x;
    if (v >= a.V) return;
    const Chunk ch = a.chunks[c];
    const int g = ch.cell;
    const float4* __restrict__ rec = a.rec + ch.start;
    const float mx = a.mu[((size_t)g * 3 + 0) * a.Vld + v];
    const float my = a.mu[((size_t)g * 3 + 1) * a.Vld + v];
    const float mz = a.mu[((size_t)g * 3 + 2) * a.Vld + v];
    const float* __restrict__ I = a.info + 9 * (size_t)g;
    const float i0 = __ldg(I + 0), i1 = __ldg(I + 1), i2 = __ldg(I + 2), i3 = __ldg(I + 3), i4 = __ldg(I + 4), i5 = __ldg(I + 5), i6 = __ldg(I + 6),
                i7 = __ldg(I + 7), i8 = __ldg(I + 8);
    const float wk = __ldg(a.w + g);
    double acc = 0.0;
    int tprev = -2;
    float4 m0, m1, m2;
    m0 = m1 = m2 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
    for (int j = 0; j < ch.count; ++j) {
        const float4 r = __ldg(rec + j);
        const int t = __float_as_int(r.w);
        float X = r.x, Y = r.y, Z = r.z;
        if (t >= 0) {
            if (t != tprev) {
                load_tform(a.Mtab, t, a.Vld, v, m0, m1, m2);
                tprev = t;
            }
            xform(m0, m1, m2, r, X, Y, Z);
        }
        const float d0 = fsub_(X, mx), d1 = fsub_(Y, my), d2 = fsub_(Z, mz);
        const float t0 = fmul_(wk, d0), t1 = fmul_(wk, d1), t2 = fmul_(wk, d2);
        const float r0 = fadd_(fmul_(t0, i0), fadd_(fmul_(t1, i3), fmul_(t2, i6)));
        const float r1 = fadd_(fmul_(t0, i1), fadd_(fmul_(t1, i4), fmul_(t2, i7)));
        const float r2 = fadd_(fmul_(t0, i2), fadd_(fmul_(t1, i5), fmul_(t2, i8)));
        const float s = fadd_(fmul_(r0, d0), fadd_(fmul_(r1, d1), fmul_(r2, d2)));
        acc += (double)s;
    }
    a.Q[(size_t)c * a.Vld + v] = acc;
}

// per (set, v): e = sqrt(|sum of chunk partials|); rows of sets owned by other ranks are zero   DmsaOptimizer.h:267
__global__ void k_cost_fin(CostArgs a, int G) {
    const int g = blockIdx.x;
    const int v = threadIdx.x;
    if (g >= G || v >= a.Vld) return;
    double q = 0.0;
    if (v < a.V) {
        const int nc = a.nchunk[g], o = a.chunk_off[g];
        for (int c = 0; c < nc; ++c) q += a.Q[(size_t)(o + c) * a.Vld + v];
    }
    a.E[(size_t)g * a.Vld + v] = sqrt(fabs(q));
}

// per-vector cost sum_r e[r][v]^2 (line search, DmsaOptimizer.h:171): one block per v, fixed reduction order
__global__ void k_col_sumsq(const double* __restrict__ E, int R, int Vld, double* __restrict__ out) {
    __shared__ double part[256];
    const int v = blockIdx.x;
    double s = 0.0;
    for (int r = threadIdx.x; r < R; r += blockDim.x) {
        double e = E[(size_t)r * Vld + v];
        s += e * e;
    }
    part[threadIdx.x] = s;
    __syncthreads();
    for (int o = blockDim.x / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) part[threadIdx.x] += part[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[v] = part[0];
}

// ---- [J e0]^T [J e0] : H = J^T J, g = J^T e0, err0 = e0^T e0 in one symmetric product ---------------------------
// Column c < P of the augmented matrix is the forward-difference column (E[:,c+1] - E[:,0]) / h (DmsaOptimizer.h:227),
// column P is e0.  Upper-triangular 32x32 tiles, split over row ranges; partials are reduced in fixed order.
#define JTJ_T 32
__global__ void __launch_bounds__(256) k_jtj(const double* __restrict__ E, int R, int Vld, int P, double inv_h, int rows_per_split,
                                             double* __restrict__ part /*[split][(P+1)*(P+1)]*/) {
    __shared__ double A[JTJ_T][JTJ_T + 1], B[JTJ_T][JTJ_T + 1];
    const int nt = (P + 1 + JTJ_T - 1) / JTJ_T;
    // decode upper-triangular tile index
    int t = blockIdx.x, ta = 0;
    while (t >= nt - ta) {
        t -= nt - ta;
        ++ta;
    }
    const int tb = ta + t;
    const int split = blockIdx.y;
    const int r0 = split * rows_per_split, r1 = min(R, r0 + rows_per_split);
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    double acc00 = 0, acc01 = 0, acc10 = 0, acc11 = 0;
    const int n1 = P + 1;
    for (int rb = r0; rb < r1; rb += JTJ_T) {
        // load 32 rows x 32 cols of each tile (256 threads, 4 elements each)
        for (int q = threadIdx.x; q < JTJ_T * JTJ_T; q += 256) {
            const int rr = q / JTJ_T, cc = q % JTJ_T;
            const int r = rb + rr;
            double va = 0.0, vb = 0.0;
            if (r < r1) {
                const double e0 = E[(size_t)r * Vld];
                const int ca = ta * JTJ_T + cc, cb = tb * JTJ_T + cc;
                if (ca < P)
                    va = inv_h * (E[(size_t)r * Vld + ca + 1] - e0);
                else if (ca == P)
                    va = e0;
                if (cb < P)
                    vb = inv_h * (E[(size_t)r * Vld + cb + 1] - e0);
                else if (cb == P)
                    vb = e0;
            }
            A[rr][cc] = va;
            B[rr][cc] = vb;
        }
        __syncthreads();
#pragma unroll 8
        for (int rr = 0; rr < JTJ_T; ++rr) {
            const double a0 = A[rr][2 * ty], a1 = A[rr][2 * ty + 1];
            const double b0 = B[rr][2 * tx], b1 = B[rr][2 * tx + 1];
            acc00 = fma(a0, b0, acc00);
            acc01 = fma(a0, b1, acc01);
            acc10 = fma(a1, b0, acc10);
            acc11 = fma(a1, b1, acc11);
        }
        __syncthreads();
    }
    double* out = part + (size_t)split * n1 * n1;
    const int i0 = ta * JTJ_T + 2 * ty, j0 = tb * JTJ_T + 2 * tx;
    if (i0 < n1 && j0 < n1) out[(size_t)i0 * n1 + j0] = acc00;
    if (i0 < n1 && j0 + 1 < n1) out[(size_t)i0 * n1 + j0 + 1] = acc01;
    if (i0 + 1 < n1 && j0 < n1) out[(size_t)(i0 + 1) * n1 + j0] = acc10;
    if (i0 + 1 < n1 && j0 + 1 < n1) out[(size_t)(i0 + 1) * n1 + j0 + 1] = acc11;
}
// out = [H (P*P row-major) | g (P) | err0]
__global__ void k_jtj_reduce(const double* __restrict__ part, int nsplit, int P, double* __restrict__ out) {
    const int n1 = P + 1;
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n1 * n1) return;
    int i = q / n1, j = q % n1;
    int ii = i, jj = j;
    if (i / JTJ_T > j / JTJ_T) {  // lower tiles were not computed: mirror
        ii = j;
        jj = i;
    }
    double s = 0.0;
    for (int k = 0; k < nsplit; ++k) s += part[(size_t)k * n1 * n1 + (size_t)ii * n1 + jj];
    if (i < P && j < P)
        out[(size_t)i * P + j] = s;
    else if (i < P && j == P)
        out[(size_t)P * P + i] = s;
    else if (i == P && j == P)
        out[(size_t)P * P + P] = s;
}

// ---- base-pose world points (updateGlobalPoints): scan points through column v of the table ----------------------
// ContinuousTrajectory.h:137-155 | MapManagement.h:140-147
__global__ void k_transform_points(const float4* __restrict__ local, const int* __restrict__ tid, int n, const float4* __restrict__ Mtab, int Vld,
                                   int v, float4* __restrict__ world, const float4* __restrict__ normal_l, float4* __restrict__ normal_w) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 p = local[i];
    const int t = tid[i];
    if (t < 0) {
        world[i] = p;
        return;
    }
    float4 m0, m1, m2;
    load_tform(Mtab, t, Vld, v, m0, m1, m2);
    float4 o;
    xform(m0, m1, m2, p, o.x, o.y, o.z);
    o.w = p.w;
    world[i] = o;
    if (normal_l) {
        // MapManagement.h:146: Matrix3f * Vector3f, coefficient redux a0 + (a1 + a2)
        const float4 nl = normal_l[i];
        float4 nw;
        nw.x = fadd_(fmul_(m0.x, nl.x), fadd_(fmul_(m0.y, nl.y), fmul_(m0.z, nl.z)));
        nw.y = fadd_(fmul_(m1.x, nl.x), fadd_(fmul_(m1.y, nl.y), fmul_(m1.z, nl.z)));
        nw.z = fadd_(fmul_(m2.x, nl.x), fadd_(fmul_(m2.y, nl.y), fmul_(m2.z, nl.z)));
        nw.w = nl.w;
        normal_w[i] = nw;
    }
}

}  // namespace dmsa
