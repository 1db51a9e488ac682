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
//   mean   = float(sum_j world_j) / float(n)        exactly-rounded sum (double accumulation; see DESIGN.md "mean")
//   d      = world - mean
//   term   = ((w*d)^T * info) * d                   row-vector * Matrix3f * vector, 3-element redux a0 + (a1 + a2)
//   e      = sqrt(|sum_j double(term_j)|)
#pragma once
#include <cuda_runtime.h>

#include "kernels_sets.cuh"

namespace dmsa {

#define COST_CHUNK 512  // members per block-sized work unit: == CHUNK == FUSE_MAX of dmsa_b200.cu (shared-memory staging buffer)

struct CostArgs {
    const Chunk* chunks;
    const int* n_chunks;     // device scalar: total chunks of the big sets (grid is an upper bound)
    const float4* rec;       // member records, sorted order: local xyz + transform-table row (int bits)
    const float4* Mtab;      // [(row * Vld + v) * 3 + r]; the last row is the identity (static points)
    int V, Vld;
    int S;                   // member sub-streams per warp: 1 when V > 16 (one thread per vector), else 32 / V lane groups
    const float* info;       // [g][9]
    const float* w;          // [g]
    const int* cell_start;   // [g]
    const int* cell_n;       // [g]
    const int* cell_kind;    // [g] 0: owned by another rank, 1: small (fused kernel), 2: big (chunked kernels)
    const int* order;        // small sets in descending-size order (longest blocks first)
    const int* nchunk;       // [g]
    const int* chunk_off;    // [g]
    double* S_part;          // partial sums   [(c*3 + a) * Vld + v]
    int* done;               // [g] chunk blocks of set g that finished pass 2 (self-resetting counter)
    double* Q;               // partial quadratic forms [c * Vld + v]
    double* E;               // residuals      [g * Vld + v]
};

__device__ __forceinline__ void xform(const float4& m0, const float4& m1, const float4& m2, const float4& r, float& X, float& Y, float& Z) {
    X = fadd_(fadd_(fadd_(fmul_(m0.x, r.x), fmul_(m0.y, r.y)), fmul_(m0.z, r.z)), m0.w);
    Y = fadd_(fadd_(fadd_(fmul_(m1.x, r.x), fmul_(m1.y, r.y)), fmul_(m1.z, r.z)), m1.w);
    Z = fadd_(fadd_(fadd_(fmul_(m2.x, r.x), fmul_(m2.y, r.y)), fmul_(m2.z, r.z)), m2.w);
}

// Lane mapping.  PACKED = false (V > 16): blockDim = Vld, thread = vector, every thread walks all members.
// PACKED = true (V <= 16, the 9 line-search vectors): blockDim = 32, lane = sub * V + v; the S = 32 / V sub-streams walk
// interleaved members and are combined in fixed order (sub 0 + sub 1 + ...) with warp shuffles.
struct LaneMap {
    int v, sub, stride;
    bool active;
};
template <bool PACKED>
__device__ __forceinline__ LaneMap lane_map(const CostArgs& a) {
    LaneMap m;
    if (!PACKED) {
        m.v = threadIdx.x;
        m.sub = 0;
        m.stride = 1;
        m.active = m.v < a.V;
        if (!m.active) m.v = a.V - 1;
    } else {
        m.sub = threadIdx.x / a.V;
        m.v = threadIdx.x - m.sub * a.V;
        m.stride = a.S;
        m.active = m.sub < a.S;
    }
    return m;
}
template <bool PACKED>
__device__ __forceinline__ double combine_subs(const CostArgs& a, const LaneMap& lm, double val) {
    if (!PACKED) return val;
    double tot = 0.0;
    for (int s = 0; s < a.S; ++s) tot += __shfl_sync(0xffffffffu, val, lm.v + s * a.V);
    return tot;
}

// Staging of `count` member records (16 B each, contiguous in HBM) into shared memory with ONE bulk asynchronous copy
// (TMA 1-D bulk copy: cp.async.bulk.shared::cluster.global + mbarrier complete_tx; SASS: UBLKCP / SYNCS): a single
// elected thread issues the copy, every thread waits on the mbarrier's phase 0.  The barrier is used once per block.
__device__ __forceinline__ void stage_records(float4* __restrict__ srec, const float4* __restrict__ rec, int count, unsigned long long* bar) {
    const unsigned bar_s = (unsigned)__cvta_generic_to_shared(bar);
    const unsigned dst_s = (unsigned)__cvta_generic_to_shared(srec);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_s));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned bytes = (unsigned)count * 16u;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_s), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_s), "l"(rec), "r"(bytes), "r"(bar_s)
                     : "memory");
    }
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "DMSA_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n"
        "@p bra DMSA_DONE;\n"
        "bra DMSA_WAIT;\n"
        "DMSA_DONE:\n"
        "}\n" ::"r"(bar_s)
        : "memory");
}

// The transform of row t for this thread's vector; fetched only when the row changes between consecutive members
// (predicated loads, no branch: the loop stays straight-line so the compiler can pipeline it).
#define DMSA_ROW_UPDATE(t)                                                                                              \
    if ((t) != tprev) {                                                                                                \
        const float4* Mp = reinterpret_cast<const float4*>(Mv + (size_t)(unsigned)(t) * (size_t)rowbytes);             \
        m0 = __ldg(Mp);                                                                                                \
        m1 = __ldg(Mp + 1);                                                                                            \
        m2 = __ldg(Mp + 2);                                                                                            \
        tprev = (t);                                                                                                   \
    }

// sum of the transformed coordinates of the staged members, sub-stream lm.sub
template <bool PACKED>
__device__ __forceinline__ void pass_sum(const CostArgs& a, const float4* __restrict__ srec, int count, const LaneMap& lm, double& sx, double& sy,
                                         double& sz) {
    sx = sy = sz = 0.0;
    int tprev = -1;
    float4 m0, m1, m2;
    m0 = m1 = m2 = make_float4(0.f, 0.f, 0.f, 0.f);
    const char* __restrict__ Mv = reinterpret_cast<const char*>(a.Mtab) + (size_t)lm.v * 48u;  // this vector's column of the table
    const unsigned rowbytes = (unsigned)a.Vld * 48u;
    const int stride = PACKED ? lm.stride : 1;
    const int nmine = (count - lm.sub + stride - 1) / stride;  // members of this sub-stream
#define DMSA_SUM_BODY(jj)                          \
    {                                              \
        const float4 r = srec[(jj)];               \
        const int t = __float_as_int(r.w);         \
        DMSA_ROW_UPDATE(t)                         \
        float X, Y, Z;                             \
        xform(m0, m1, m2, r, X, Y, Z);             \
        sx += (double)X;                           \
        sy += (double)Y;                           \
        sz += (double)Z;                           \
    }
    int k = 0, j = lm.sub;
    for (; k + 4 <= nmine; k += 4, j += 4 * stride) {  // no exit test inside a group of four
        DMSA_SUM_BODY(j)
        DMSA_SUM_BODY(j + stride)
        DMSA_SUM_BODY(j + 2 * stride)
        DMSA_SUM_BODY(j + 3 * stride)
    }
    for (; k < nmine; ++k, j += stride) DMSA_SUM_BODY(j)
#undef DMSA_SUM_BODY
}
// sum of the Mahalanobis terms  ((w d)^T info) d                                   DmsaOptimizer.h:259-264
template <bool PACKED>
__device__ __forceinline__ double pass_quad(const CostArgs& a, const float4* __restrict__ srec, int count, const LaneMap& lm, int g, float mx, float my,
                                            float mz) {
    const float* __restrict__ I = a.info + 9 * (size_t)g;
    const float i0 = __ldg(I + 0), i1 = __ldg(I + 1), i2 = __ldg(I + 2), i3 = __ldg(I + 3), i4 = __ldg(I + 4), i5 = __ldg(I + 5), i6 = __ldg(I + 6),
                i7 = __ldg(I + 7), i8 = __ldg(I + 8);
    const float wk = __ldg(a.w + g);
    double acc = 0.0;
    int tprev = -1;
    float4 m0, m1, m2;
    m0 = m1 = m2 = make_float4(0.f, 0.f, 0.f, 0.f);
    const char* __restrict__ Mv = reinterpret_cast<const char*>(a.Mtab) + (size_t)lm.v * 48u;  // this vector's column of the table
    const unsigned rowbytes = (unsigned)a.Vld * 48u;
    const int stride = PACKED ? lm.stride : 1;
    const int nmine = (count - lm.sub + stride - 1) / stride;
#define DMSA_QUAD_BODY(jj)                                                                   \
    {                                                                                        \
        const float4 r = srec[(jj)];                                                         \
        const int t = __float_as_int(r.w);                                                   \
        DMSA_ROW_UPDATE(t)                                                                   \
        float X, Y, Z;                                                                       \
        xform(m0, m1, m2, r, X, Y, Z);                                                       \
        const float d0 = fsub_(X, mx), d1 = fsub_(Y, my), d2 = fsub_(Z, mz);                 \
        const float t0 = fmul_(wk, d0), t1 = fmul_(wk, d1), t2 = fmul_(wk, d2);              \
        const float r0 = fadd_(fmul_(t0, i0), fadd_(fmul_(t1, i3), fmul_(t2, i6)));          \
        const float r1 = fadd_(fmul_(t0, i1), fadd_(fmul_(t1, i4), fmul_(t2, i7)));          \
        const float r2 = fadd_(fmul_(t0, i2), fadd_(fmul_(t1, i5), fmul_(t2, i8)));          \
        const float s_ = fadd_(fmul_(r0, d0), fadd_(fmul_(r1, d1), fmul_(r2, d2)));          \
        acc += (double)s_;                                                                   \
    }
    int k = 0, j = lm.sub;
    for (; k + 4 <= nmine; k += 4, j += 4 * stride) {
        DMSA_QUAD_BODY(j)
        DMSA_QUAD_BODY(j + stride)
        DMSA_QUAD_BODY(j + 2 * stride)
        DMSA_QUAD_BODY(j + 3 * stride)
    }
    for (; k < nmine; ++k, j += stride) DMSA_QUAD_BODY(j)
#undef DMSA_QUAD_BODY
    return acc;
}

// Small sets (n <= COST_CHUNK): one block stages the set's member records in shared memory once and runs both passes
// on them.  Blocks are issued longest-set-first (a.order) so that the kernel does not end on a long block.
// Also zero-fills the rows of sets owned by other ranks.
template <bool PACKED, int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB) k_cost_fused(CostArgs a, int G) {
    __shared__ __align__(128) float4 srec[COST_CHUNK];
    __shared__ __align__(8) unsigned long long bar;
    if ((int)blockIdx.x >= G) return;
    const int g = a.order[blockIdx.x];
    const int kind = a.cell_kind[g];
    if (kind == 2) return;
    const LaneMap lm = lane_map<PACKED>(a);
    if (kind == 0) {
        if (lm.active && lm.sub == 0) a.E[(size_t)g * a.Vld + lm.v] = 0.0;
        return;
    }
    const int n = a.cell_n[g];
    stage_records(srec, a.rec + a.cell_start[g], n, &bar);
    const int cnt = lm.active ? n : 0;
    double sx, sy, sz;
    pass_sum<PACKED>(a, srec, cnt, lm, sx, sy, sz);
    sx = combine_subs<PACKED>(a, lm, sx);
    sy = combine_subs<PACKED>(a, lm, sy);
    sz = combine_subs<PACKED>(a, lm, sz);
    const float nf = (float)n;
    const float mx = fdiv_((float)sx, nf), my = fdiv_((float)sy, nf), mz = fdiv_((float)sz, nf);  // DmsaOptimizer.h:254
    double q = pass_quad<PACKED>(a, srec, cnt, lm, g, mx, my, mz);
    q = combine_subs<PACKED>(a, lm, q);
    if (lm.active && lm.sub == 0) a.E[(size_t)g * a.Vld + lm.v] = sqrt(fabs(q));  // :267
}

// Reference-order validation kernel: the per-set mean accumulated SEQUENTIALLY IN FLOAT in member order, exactly like
// DmsaOptimizer.h:249-254 (`mean = mean + x_j`), for every set regardless of size (one block per set, members staged
// in tiles).  Slow on the big sets of un-downsampled clouds (a dependent float-add chain as long as the set) and
// therefore not the default; it exists to show that the only arithmetic difference between the fast path and the
// reference's operation order is that one reduction (DESIGN.md §3 "mean").
__global__ void __launch_bounds__(1024) k_cost_seq(CostArgs a, int G) {
    __shared__ __align__(16) float4 srec[COST_CHUNK];
    const int g = blockIdx.x;
    if (g >= G) return;
    const int kind = a.cell_kind[g];
    const int v = min((int)threadIdx.x, a.V - 1);
    const bool active = (int)threadIdx.x < a.V;
    if (kind == 0) {
        if (active) a.E[(size_t)g * a.Vld + v] = 0.0;
        return;
    }
    const int n = a.cell_n[g];
    const float4* __restrict__ rec = a.rec + a.cell_start[g];
    const char* __restrict__ Mv = reinterpret_cast<const char*>(a.Mtab) + (size_t)v * 48u;
    const unsigned rowbytes = (unsigned)a.Vld * 48u;
    int tprev = -1;
    float4 m0, m1, m2;
    m0 = m1 = m2 = make_float4(0.f, 0.f, 0.f, 0.f);
    float sx = 0.f, sy = 0.f, sz = 0.f;
    for (int base = 0; base < n; base += COST_CHUNK) {
        const int cnt = min(COST_CHUNK, n - base);
        __syncthreads();
        for (int i = threadIdx.x; i < cnt; i += blockDim.x) srec[i] = __ldg(rec + base + i);
        __syncthreads();
        for (int j = 0; j < cnt; ++j) {
            const float4 r = srec[j];
            const int t = __float_as_int(r.w);
            DMSA_ROW_UPDATE(t)
            float X, Y, Z;
            xform(m0, m1, m2, r, X, Y, Z);
            sx = fadd_(sx, X);  // :251 sequential float accumulation
            sy = fadd_(sy, Y);
            sz = fadd_(sz, Z);
        }
    }
    const float nf = (float)n;
    const float mx = fdiv_(sx, nf), my = fdiv_(sy, nf), mz = fdiv_(sz, nf);  // :254
    const float* __restrict__ I = a.info + 9 * (size_t)g;
    const float i0 = I[0], i1 = I[1], i2 = I[2], i3 = I[3], i4 = I[4], i5 = I[5], i6 = I[6], i7 = I[7], i8 = I[8];
    const float wk = a.w[g];
    double acc = 0.0;
    for (int base = 0; base < n; base += COST_CHUNK) {
        const int cnt = min(COST_CHUNK, n - base);
        __syncthreads();
        for (int i = threadIdx.x; i < cnt; i += blockDim.x) srec[i] = __ldg(rec + base + i);
        __syncthreads();
        for (int j = 0; j < cnt; ++j) {
            const float4 r = srec[j];
            const int t = __float_as_int(r.w);
            DMSA_ROW_UPDATE(t)
            float X, Y, Z;
            xform(m0, m1, m2, r, X, Y, Z);
            const float d0 = fsub_(X, mx), d1 = fsub_(Y, my), d2 = fsub_(Z, mz);
            const float t0 = fmul_(wk, d0), t1 = fmul_(wk, d1), t2 = fmul_(wk, d2);
            const float r0 = fadd_(fmul_(t0, i0), fadd_(fmul_(t1, i3), fmul_(t2, i6)));
            const float r1 = fadd_(fmul_(t0, i1), fadd_(fmul_(t1, i4), fmul_(t2, i7)));
            const float r2 = fadd_(fmul_(t0, i2), fadd_(fmul_(t1, i5), fmul_(t2, i8)));
            const float s_ = fadd_(fmul_(r0, d0), fadd_(fmul_(r1, d1), fmul_(r2, d2)));
            acc += (double)s_;  // :263 errorVec(k) += float term, in member order
        }
    }
    if (active) a.E[(size_t)g * a.Vld + v] = sqrt(fabs(acc));
}

// Big sets, pass 1: per-chunk coordinate sums (double)
template <bool PACKED, int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB) k_cost_sum(CostArgs a) {
    __shared__ __align__(128) float4 srec[COST_CHUNK];
    __shared__ __align__(8) unsigned long long bar;
    const int c = blockIdx.x;
    if (c >= *a.n_chunks) return;
    const LaneMap lm = lane_map<PACKED>(a);
    const Chunk ch = a.chunks[c];
    stage_records(srec, a.rec + ch.start, ch.count, &bar);
    double sx, sy, sz;
    pass_sum<PACKED>(a, srec, lm.active ? ch.count : 0, lm, sx, sy, sz);
    sx = combine_subs<PACKED>(a, lm, sx);
    sy = combine_subs<PACKED>(a, lm, sy);
    sz = combine_subs<PACKED>(a, lm, sz);
    if (lm.active && lm.sub == 0) {
        a.S_part[((size_t)c * 3 + 0) * a.Vld + lm.v] = sx;
        a.S_part[((size_t)c * 3 + 1) * a.Vld + lm.v] = sy;
        a.S_part[((size_t)c * 3 + 2) * a.Vld + lm.v] = sz;
    }
}

#define COST_RED_Y 8
// Fixed-order reduction of a big set's chunk partials, identical for every reader: COST_RED_Y interleaved partial sums
// (chunk c goes to slot c % COST_RED_Y, ascending c), then the slots in ascending order.  Every block of the set
// recomputes it from the same partials, so all of them see bit-identical means.
__device__ __forceinline__ double reduce_chunks(const double* __restrict__ part, int nc, size_t stride) {
    double t = 0.0;
#pragma unroll 1
    for (int y = 0; y < COST_RED_Y; ++y) {
        double s = 0.0;
        for (int c = y; c < nc; c += COST_RED_Y) s += __ldcg(part + (size_t)c * stride);
        t += s;
    }
    return t;
}

// Big sets, pass 2: mean = float(sum over the set's chunks) / float(n) (DmsaOptimizer.h:254, recomputed by every chunk
// block from the chunk partials of pass 1), per-chunk sums of the Mahalanobis terms, and - in the block that finishes
// last for its set (fence + per-set counter) - e = sqrt(|sum of the chunk partials|) in fixed chunk order (:267).
template <bool PACKED, int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB) k_cost_quad(CostArgs a) {
    __shared__ __align__(128) float4 srec[COST_CHUNK];
    __shared__ __align__(8) unsigned long long bar;
    __shared__ int s_last;
    const int c = blockIdx.x;
    if (c >= *a.n_chunks) return;
    const LaneMap lm = lane_map<PACKED>(a);
    const Chunk ch = a.chunks[c];
    const int g = ch.cell;
    stage_records(srec, a.rec + ch.start, ch.count, &bar);
    const int nc = a.nchunk[g], o = ch.first;
    const float nf = (float)a.cell_n[g];
    const size_t st3 = (size_t)3 * a.Vld;
    const double* __restrict__ Sp = a.S_part + (size_t)o * st3 + lm.v;
    const float mx = fdiv_((float)reduce_chunks(Sp, nc, st3), nf);
    const float my = fdiv_((float)reduce_chunks(Sp + a.Vld, nc, st3), nf);
    const float mz = fdiv_((float)reduce_chunks(Sp + 2 * (size_t)a.Vld, nc, st3), nf);
    double acc = pass_quad<PACKED>(a, srec, lm.active ? ch.count : 0, lm, g, mx, my, mz);
    acc = combine_subs<PACKED>(a, lm, acc);
    const bool writer = lm.active && lm.sub == 0;
    if (writer) a.Q[(size_t)c * a.Vld + lm.v] = acc;
    __threadfence();  // the partials of this block are visible device-wide before the counter moves
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(a.done + g, 1) == nc - 1) ? 1 : 0;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    if (writer) a.E[(size_t)g * a.Vld + lm.v] = sqrt(fabs(reduce_chunks(a.Q + (size_t)o * a.Vld + lm.v, nc, (size_t)a.Vld)));
    if (threadIdx.x == 0) a.done[g] = 0;  // ready for the next launch
}

// per-vector cost sum_r e[r][v]^2 (line search, DmsaOptimizer.h:171): COLSUM_PARTS row slices per vector, fixed reduction order
#define COLSUM_PARTS 16
__global__ void k_col_sumsq(const double* __restrict__ E, int R, int Vld, double* __restrict__ part /*[9][COLSUM_PARTS]*/) {
    __shared__ double red[256];
    const int v = blockIdx.x, y = blockIdx.y;
    const int per = (R + COLSUM_PARTS - 1) / COLSUM_PARTS;
    const int r0 = y * per, r1 = min(R, r0 + per);
    double s = 0.0;
    for (int r = r0 + threadIdx.x; r < r1; r += blockDim.x) {
        double e = E[(size_t)r * Vld + v];
        s += e * e;
    }
    red[threadIdx.x] = s;
    __syncthreads();
    for (int o = blockDim.x / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) part[v * COLSUM_PARTS + y] = red[0];
}
__global__ void k_col_sumsq_fin(const double* __restrict__ part, double* __restrict__ out) {
    const int v = threadIdx.x;
    if (v >= 9) return;
    double s = 0.0;
    for (int y = 0; y < COLSUM_PARTS; ++y) s += part[v * COLSUM_PARTS + y];
    out[v] = s;
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
    const float4* Mp = Mtab + ((size_t)t * Vld + v) * 3;
    const float4 m0 = __ldg(Mp), m1 = __ldg(Mp + 1), m2 = __ldg(Mp + 2);
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
