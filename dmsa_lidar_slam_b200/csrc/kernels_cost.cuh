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

#include "pdl.cuh"

#include "kernels_sets.cuh"

namespace dmsa {

#define COST_CHUNK 512  // members per block-sized work unit: == CHUNK == FUSE_MAX of dmsa_b200.cu (shared-memory staging buffer)

struct CostArgs {
    const Chunk* chunks;
    const int* n_chunks;     // device scalar: total chunks of the big sets (grid is an upper bound)
    const float4* rec;       // member records, sorted order: local xyz + transform-table row (int bits)
    const float4* Mtab;      // [(row * Vld + v) * 3 + r]; the last row is the identity (static points)
    const unsigned long long* Mpair;  // pair-interleaved copy [(row * Vp + tp) * 12 + c] = {M[row][2tp][c], M[row][2tp+1][c]}
    int Vp;                  // Vld / 2
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
    float* mu;               // means of the sets with more than MEAN_INLINE_MAX chunks [(g*3 + a) * Vld + v] (written by pass 1)
    int* done;               // [g] chunk blocks of set g that finished pass 2 (self-resetting counter)
    int* done1;              // [g] same for pass 1 (sets with more than MEAN_INLINE_MAX chunks only)
    double* Q;               // partial quadratic forms [c * Vld + v]
    double* E;               // residuals      [g * Vld + v]
    const LevelInfo* li;     // the set count lives on the device (total_sets): grids are sized from a host-side bound
    // pair kernels, forward-difference batch with the shared-rotation fast path (split = 1): every work unit (set / chunk) is
    // evaluated by TWO independent blocks, blockIdx.x = 2 unit + role.  Role 0 owns the vector pairs [0, ns), which carry a
    // rotation of their own; role 1 owns the pairs [ns, ..), which perturb a translation only and share the rotated member
    // coordinates of vector 0.  split = 0: one block per unit evaluates every pair in full.
    int split, ns;
    int* done_f;             // completion counters of role 1 (k_cost_quad2) [g]
    int* done1_f;            // ... and of k_cost_sum2
};

__device__ __forceinline__ void xform(const float4& m0, const float4& m1, const float4& m2, const float4& r, float& X, float& Y, float& Z) {
    X = fadd_(fadd_(fadd_(fmul_(m0.x, r.x), fmul_(m0.y, r.y)), fmul_(m0.z, r.z)), m0.w);
    Y = fadd_(fadd_(fadd_(fmul_(m1.x, r.x), fmul_(m1.y, r.y)), fmul_(m1.z, r.z)), m1.w);
    Z = fadd_(fadd_(fadd_(fmul_(m2.x, r.x), fmul_(m2.y, r.y)), fmul_(m2.z, r.z)), m2.w);
}

// Lane mapping.  PACKED = false (V > 16): blockDim = Vld, thread = vector, every thread walks all members.
// PACKED = true (V <= 16, the 9 line-search vectors): blockDim = 32 * PACKED_WARPS; in every warp lane = sub * V + v, the
// S = 32 / V sub-streams of each warp walk interleaved members (PACKED_WARPS * S streams per set) and are combined in fixed
// order: inside a warp with shuffles (sub 0 + sub 1 + ...), then warp 0 + warp 1 through shared memory.  The small batch is
// latency bound (one dependent chain per lane): two warps per set double the resident warps for the same staging buffer.
#define PACKED_WARPS 2
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
        const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
        const int sw = lane / a.V;  // sub-stream inside the warp
        m.v = lane - sw * a.V;
        m.sub = w * a.S + sw;
        m.stride = a.S * PACKED_WARPS;
        m.active = sw < a.S;
    }
    return m;
}
// ex: shared double[PACKED_WARPS][16].  Every thread returns the total of its vector (block-uniform per vector).
template <bool PACKED>
__device__ __forceinline__ double combine_subs(const CostArgs& a, const LaneMap& lm, double val, double (*ex)[16]) {
    if (!PACKED) return val;
    double tot = 0.0;
    for (int s = 0; s < a.S; ++s) tot += __shfl_sync(0xffffffffu, val, lm.v + s * a.V);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane < a.V) ex[w][lane] = tot;
    __syncthreads();
    double all = ex[0][lm.v];
#pragma unroll
    for (int k = 1; k < PACKED_WARPS; ++k) all += ex[k][lm.v];
    __syncthreads();  // ex is reused by the next combine
    return all;
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
__global__ void __launch_bounds__(MAXT, MINB) k_cost_fused(CostArgs a) {
    DMSA_PDL_ENTER();
    __shared__ __align__(128) float4 srec[COST_CHUNK];
    __shared__ __align__(8) unsigned long long bar;
    __shared__ double ex[PACKED_WARPS][16];
    if ((int)blockIdx.x >= total_sets(a.li)) return;
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
    sx = combine_subs<PACKED>(a, lm, sx, ex);
    sy = combine_subs<PACKED>(a, lm, sy, ex);
    sz = combine_subs<PACKED>(a, lm, sz, ex);
    const float nf = (float)n;
    const float mx = fdiv_((float)sx, nf), my = fdiv_((float)sy, nf), mz = fdiv_((float)sz, nf);  // DmsaOptimizer.h:254
    double q = pass_quad<PACKED>(a, srec, cnt, lm, g, mx, my, mz);
    q = combine_subs<PACKED>(a, lm, q, ex);
    if (lm.active && lm.sub == 0) a.E[(size_t)g * a.Vld + lm.v] = sqrt(fabs(q));  // :267
}

// Reference-order validation kernel: the per-set mean accumulated SEQUENTIALLY IN FLOAT in member order, exactly like
// DmsaOptimizer.h:249-254 (`mean = mean + x_j`), for every set regardless of size (one block per set, members staged
// in tiles).  Slow on the big sets of un-downsampled clouds (a dependent float-add chain as long as the set) and
// therefore not the default; it exists to show that the only arithmetic difference between the fast path and the
// reference's operation order is that one reduction (DESIGN.md §3 "mean").
__global__ void __launch_bounds__(1024) k_cost_seq(CostArgs a) {
    DMSA_PDL_ENTER();
    __shared__ __align__(16) float4 srec[COST_CHUNK];
    const int g = blockIdx.x;
    if (g >= total_sets(a.li)) return;
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

#define COST_RED_Y 8
#define MEAN_INLINE_MAX 128  // sets with at most this many chunks (65 536 members): pass 2 recomputes the mean per chunk block (measured
                             // cheaper than a hand-over); beyond it the redundant reads would grow quadratically, pass 1 hands the mean over
// Fixed-order reduction of a big set's chunk partials: COST_RED_Y interleaved partial sums (chunk c goes to stream
// c % COST_RED_Y, ascending c), then the streams in ascending order.
__device__ __forceinline__ double reduce_chunks(const double* __restrict__ part, int nc, size_t stride) {
    double s[COST_RED_Y];
#pragma unroll
    for (int y = 0; y < COST_RED_Y; ++y) s[y] = 0.0;
    int c = 0;
    for (; c + COST_RED_Y <= nc; c += COST_RED_Y) {  // eight independent loads in flight, one per stream
#pragma unroll
        for (int y = 0; y < COST_RED_Y; ++y) s[y] += __ldcg(part + (size_t)(c + y) * stride);
    }
#pragma unroll
    for (int y = 0; y < COST_RED_Y; ++y)
        if (c + y < nc) s[y] += __ldcg(part + (size_t)(c + y) * stride);
    double t = 0.0;
#pragma unroll
    for (int y = 0; y < COST_RED_Y; ++y) t += s[y];
    return t;
}
// the same reduction for two adjacent vectors at once (16-byte loads; identical order per component)
__device__ __forceinline__ double2 reduce_chunks2(const double* __restrict__ part, int nc, size_t stride) {
    double2 t = make_double2(0.0, 0.0);
#pragma unroll 1
    for (int h = 0; h < COST_RED_Y; h += 4) {  // four streams (loads in flight) at a time: half the registers of eight
        double2 s[4];
#pragma unroll
        for (int y = 0; y < 4; ++y) s[y] = make_double2(0.0, 0.0);
        int c = 0;
        for (; c + COST_RED_Y <= nc; c += COST_RED_Y) {
#pragma unroll
            for (int y = 0; y < 4; ++y) {
                const double2 q = __ldcg(reinterpret_cast<const double2*>(part + (size_t)(c + h + y) * stride));
                s[y].x += q.x;
                s[y].y += q.y;
            }
        }
#pragma unroll
        for (int y = 0; y < 4; ++y)
            if (c + h + y < nc) {
                const double2 q = __ldcg(reinterpret_cast<const double2*>(part + (size_t)(c + h + y) * stride));
                s[y].x += q.x;
                s[y].y += q.y;
            }
#pragma unroll
        for (int y = 0; y < 4; ++y) {
            t.x += s[y].x;
            t.y += s[y].y;
        }
    }
    return t;
}
// "last block of the set": fence, count, and tell the whole block whether every other chunk block of set g is done
__device__ __forceinline__ bool last_block_of_set(int* counter, int nc, int* s_flag) {
    __threadfence();  // this block's partials are visible device-wide before the counter moves
    __syncthreads();
    if (threadIdx.x == 0) *s_flag = (atomicAdd(counter, 1) == nc - 1) ? 1 : 0;
    __syncthreads();
    if (!*s_flag) return false;
    __threadfence();
    return true;
}

// Big sets, pass 1: per-chunk coordinate sums (double)
template <bool PACKED, int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB) k_cost_sum(CostArgs a) {
    DMSA_PDL_ENTER();
    __shared__ __align__(128) float4 srec[COST_CHUNK];
    __shared__ __align__(8) unsigned long long bar;
    __shared__ double ex[PACKED_WARPS][16];
    __shared__ int s_last;
    const int c = blockIdx.x;
    if (c >= *a.n_chunks) return;
    const LaneMap lm = lane_map<PACKED>(a);
    const Chunk ch = a.chunks[c];
    stage_records(srec, a.rec + ch.start, ch.count, &bar);
    double sx, sy, sz;
    pass_sum<PACKED>(a, srec, lm.active ? ch.count : 0, lm, sx, sy, sz);
    sx = combine_subs<PACKED>(a, lm, sx, ex);
    sy = combine_subs<PACKED>(a, lm, sy, ex);
    sz = combine_subs<PACKED>(a, lm, sz, ex);
    const bool writer = lm.active && lm.sub == 0;
    if (writer) {
        a.S_part[((size_t)c * 3 + 0) * a.Vld + lm.v] = sx;
        a.S_part[((size_t)c * 3 + 1) * a.Vld + lm.v] = sy;
        a.S_part[((size_t)c * 3 + 2) * a.Vld + lm.v] = sz;
    }
    // sets cut into many chunks: the block that finishes last reduces the partials to the set's mean once (pass 2 would
    // otherwise repeat that reduction in every one of the set's chunk blocks)
    const int g = ch.cell, nc = a.nchunk[g];
    if (nc <= MEAN_INLINE_MAX) return;
    if (!last_block_of_set(a.done1 + g, nc, &s_last)) return;
    if (writer) {
        const float nf = (float)a.cell_n[g];
        const size_t st3 = (size_t)3 * a.Vld;
        const double* __restrict__ Sp = a.S_part + (size_t)ch.first * st3 + lm.v;
        float* __restrict__ mu = a.mu + (size_t)g * st3 + lm.v;
        mu[0] = fdiv_((float)reduce_chunks(Sp, nc, st3), nf);  // DmsaOptimizer.h:254
        mu[a.Vld] = fdiv_((float)reduce_chunks(Sp + a.Vld, nc, st3), nf);
        mu[2 * (size_t)a.Vld] = fdiv_((float)reduce_chunks(Sp + 2 * (size_t)a.Vld, nc, st3), nf);
    }
    if (threadIdx.x == 0) a.done1[g] = 0;  // ready for the next launch
}

// Big sets, pass 2: mean = float(sum over the set's chunks) / float(n) (DmsaOptimizer.h:254; every chunk block of the set
// recomputes it from the same pass-1 partials in the same fixed order, so all of them see bit-identical means - measured
// cheaper than a separate mean kernel or a last-block reduction in pass 1), per-chunk sums of the Mahalanobis terms, and -
// in the block that finishes last for its set - e = sqrt(|sum of the chunk partials|) in fixed chunk order (:267).
template <bool PACKED, int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB) k_cost_quad(CostArgs a) {
    DMSA_PDL_ENTER();
    __shared__ __align__(128) float4 srec[COST_CHUNK];
    __shared__ __align__(8) unsigned long long bar;
    __shared__ double ex[PACKED_WARPS][16];
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
    float mx, my, mz;
    if (nc <= MEAN_INLINE_MAX) {
        const double* __restrict__ Sp = a.S_part + (size_t)o * st3 + lm.v;
        mx = fdiv_((float)reduce_chunks(Sp, nc, st3), nf);
        my = fdiv_((float)reduce_chunks(Sp + a.Vld, nc, st3), nf);
        mz = fdiv_((float)reduce_chunks(Sp + 2 * (size_t)a.Vld, nc, st3), nf);
    } else {
        const float* __restrict__ mu = a.mu + (size_t)g * st3 + lm.v;
        mx = mu[0];
        my = mu[a.Vld];
        mz = mu[2 * (size_t)a.Vld];
    }
    double acc = pass_quad<PACKED>(a, srec, lm.active ? ch.count : 0, lm, g, mx, my, mz);
    acc = combine_subs<PACKED>(a, lm, acc, ex);
    const bool writer = lm.active && lm.sub == 0;
    if (writer) a.Q[(size_t)c * a.Vld + lm.v] = acc;
    if (!last_block_of_set(a.done + g, nc, &s_last)) return;
    if (writer) a.E[(size_t)g * a.Vld + lm.v] = sqrt(fabs(reduce_chunks(a.Q + (size_t)o * a.Vld + lm.v, nc, (size_t)a.Vld)));
    if (threadIdx.x == 0) a.done[g] = 0;  // ready for the next launch
}

// =====================================================================================================================
// Pair-packed variants of the three cost kernels for the forward-difference batch (V > 16): one thread evaluates TWO
// parameter vectors (2 tp, 2 tp + 1) with Blackwell's packed FP32x2 instructions (SASS FMUL2 / FADD2).
//
// Why not simply pack everything: ptxas 12.9 contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 even under --fmad=false
// (scripts/microbench/packed_fp32.cu shows the SASS and the changed results), and the reference arithmetic has no FMA.
// The faithful packing keeps one SCALAR side on every multiply -> add edge: products are packed (FMUL2 with the member
// coordinate / information entry as broadcast scalar operand), the adds they feed are scalar FADDs on the two halves,
// adds that follow adds (+ translation, - mean) are packed again.  Per member and vector pair: 42 packed + 40 scalar
// FP32 instructions instead of 124 scalar ones, the shared-memory load, the row test and the loop are shared by the
// two vectors.  Bit-identical to the scalar kernels (tests/test_gpu_parity.py::test_pair_packed_...).
// The transform table is read from a pair-interleaved copy: Mpair[(row * Vp + tp) * 12 + c] = {M[row][2tp][c], M[row][2tp+1][c]}.
// =====================================================================================================================
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk2(float lo, float hi) {
    u64 r;
    asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void upk2(u64 p, float& lo, float& hi) { asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(p)); }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) {
    u64 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ u64 add2(u64 a, u64 b) {
    u64 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ u64 sub2(u64 a, u64 b) {
    u64 r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ u64 mul2s(u64 a, float s) { return mul2(a, pk2(s, s)); }

#define DMSA_ROW_UPDATE2(t)                                                                                              \
    if ((t) != tprev) {                                                                                                 \
        const ulonglong2* Mp = reinterpret_cast<const ulonglong2*>(Mv + (size_t)(unsigned)(t) * (size_t)rowbytes);      \
        _Pragma("unroll") for (int c_ = 0; c_ < 6; ++c_) {                                                              \
            const ulonglong2 q_ = __ldg(Mp + c_);                                                                       \
            m[2 * c_] = q_.x;                                                                                           \
            m[2 * c_ + 1] = q_.y;                                                                                       \
        }                                                                                                               \
        tprev = (t);                                                                                                    \
    }
// one row of Matrix4f * Vector4f for the vector pair: ((m0 x + m1 y) + m2 z) + m3, products packed, their adds scalar
#define DMSA_XROW2(b, r, OUT)                                                                                            \
    {                                                                                                                   \
        const u64 p0_ = mul2s(m[(b)], (r).x), p1_ = mul2s(m[(b) + 1], (r).y), p2_ = mul2s(m[(b) + 2], (r).z);           \
        float p0a, p0b, p1a, p1b, p2a, p2b;                                                                             \
        upk2(p0_, p0a, p0b);                                                                                            \
        upk2(p1_, p1a, p1b);                                                                                            \
        upk2(p2_, p2a, p2b);                                                                                            \
        const float qa_ = fadd_(fadd_(p0a, p1a), p2a);                                                                  \
        const float qb_ = fadd_(fadd_(p0b, p1b), p2b);                                                                  \
        OUT = add2(pk2(qa_, qb_), m[(b) + 3]);                                                                          \
    }
// a0 + (a1 + a2) on both halves of three packed products (Eigen's 3-element redux order)
#define DMSA_DOT3_2(P0, P1, P2, OUTA, OUTB)                                                                              \
    {                                                                                                                   \
        float a0_, b0_, a1_, b1_, a2_, b2_;                                                                             \
        upk2(P0, a0_, b0_);                                                                                             \
        upk2(P1, a1_, b1_);                                                                                             \
        upk2(P2, a2_, b2_);                                                                                             \
        OUTA = fadd_(a0_, fadd_(a1_, a2_));                                                                             \
        OUTB = fadd_(b0_, fadd_(b1_, b2_));                                                                             \
    }

// ---- shared-rotation fast path ------------------------------------------------------------------------------------
// Half of the forward-difference vectors perturb a TRANSLATION parameter (parameter layout [w_1..w_{n-1} | t_1..t_{n-1}],
// Poses.h:64-76).  The rotation block of every table row of such a vector is bit-identical to vector 0's: the pose-chain
// and dense-table kernels compute it from the rotation parameters alone, by the same instruction sequence
// (tests/test_gpu_parity.py::test_translation_vectors_share_the_base_rotation_bitwise checks the tables).  The first three
// terms of Eigen's Matrix4f * Vector4f product, q = (m0 x + m1 y) + m2 z, are therefore the same floats for vector 0 and
// for every translation vector; a block of translation pairs computes q once per member (stage_partials) and its threads
// only add their own translation: X = q + m3 — 3 packed adds instead of 9 packed multiplies + 12 adds + 3 packed adds.
// The two kinds of pairs run in separate blocks (CostArgs::split): the kernels are bound by the latency of each warp's
// dependent chain, not by issue slots, so a block that mixes a full-cost warp with a cheap one would hold its slot on the SM
// for the full-cost duration and gain nothing (measured); separate blocks free their slots as soon as they finish.
#define DMSA_ROW_UPDATE2F(t)                                                                                             \
    if ((t) != tprev) {                                                                                                 \
        const u64* Mp = reinterpret_cast<const u64*>(Mv + (size_t)(unsigned)(t) * (size_t)rowbytes);                    \
        m[3] = __ldg(Mp + 3);                                                                                           \
        m[7] = __ldg(Mp + 7);                                                                                           \
        m[11] = __ldg(Mp + 11);                                                                                         \
        tprev = (t);                                                                                                    \
    }
// world coordinates of member record r for the vector pair (FAST: r holds the rotated coordinates q)
#define DMSA_XFORM2(r, X, Y, Z)                 \
    if constexpr (FAST) {                       \
        DMSA_ROW_UPDATE2F(t)                    \
        X = add2(pk2((r).x, (r).x), m[3]);      \
        Y = add2(pk2((r).y, (r).y), m[7]);      \
        Z = add2(pk2((r).z, (r).z), m[11]);     \
    } else {                                    \
        DMSA_ROW_UPDATE2(t)                     \
        DMSA_XROW2(0, r, X)                     \
        DMSA_XROW2(4, r, Y)                     \
        DMSA_XROW2(8, r, Z)                     \
    }
// q_j = (m0 x + m1 y) + m2 z with vector 0's rotation for `count` member records (coalesced 16-byte loads); sq[j] = (q, table row)
__device__ __forceinline__ void stage_partials(const CostArgs& a, const float4* __restrict__ rec, float4* __restrict__ sq, int count) {
    for (int j = threadIdx.x; j < count; j += blockDim.x) {
        const float4 r = __ldg(rec + j);
        const float4* __restrict__ Mp = a.Mtab + (size_t)(unsigned)__float_as_int(r.w) * (size_t)a.Vld * 3;
        const float4 m0 = __ldg(Mp), m1 = __ldg(Mp + 1), m2 = __ldg(Mp + 2);
        float4 q;
        q.x = fadd_(fadd_(fmul_(m0.x, r.x), fmul_(m0.y, r.y)), fmul_(m0.z, r.z));
        q.y = fadd_(fadd_(fmul_(m1.x, r.x), fmul_(m1.y, r.y)), fmul_(m1.z, r.z));
        q.z = fadd_(fadd_(fmul_(m2.x, r.x), fmul_(m2.y, r.y)), fmul_(m2.z, r.z));
        q.w = r.w;
        sq[j] = q;
    }
    __syncthreads();
}

struct PairSums {
    double x0, y0, z0, x1, y1, z1;
};
template <bool FAST>
__device__ __forceinline__ void pass_sum2(const CostArgs& a, const float4* __restrict__ srec, int count, int tp, PairSums& S) {
    S.x0 = S.y0 = S.z0 = S.x1 = S.y1 = S.z1 = 0.0;
    int tprev = -1;
    u64 m[12];
#pragma unroll
    for (int c = 0; c < 12; ++c) m[c] = 0ull;
    const char* __restrict__ Mv = reinterpret_cast<const char*>(a.Mpair) + (size_t)tp * 96u;
    const unsigned rowbytes = (unsigned)a.Vp * 96u;
#define DMSA_SUM_BODY2(jj)                   \
    {                                        \
        const float4 r = srec[(jj)];         \
        const int t = __float_as_int(r.w);   \
        u64 X, Y, Z;                         \
        DMSA_XFORM2(r, X, Y, Z)              \
        float lo, hi;                        \
        upk2(X, lo, hi);                     \
        S.x0 += (double)lo;                  \
        S.x1 += (double)hi;                  \
        upk2(Y, lo, hi);                     \
        S.y0 += (double)lo;                  \
        S.y1 += (double)hi;                  \
        upk2(Z, lo, hi);                     \
        S.z0 += (double)lo;                  \
        S.z1 += (double)hi;                  \
    }
    int j = 0;
    for (; j + 4 <= count; j += 4) {
        DMSA_SUM_BODY2(j)
        DMSA_SUM_BODY2(j + 1)
        DMSA_SUM_BODY2(j + 2)
        DMSA_SUM_BODY2(j + 3)
    }
    for (; j < count; ++j) DMSA_SUM_BODY2(j)
#undef DMSA_SUM_BODY2
}
template <bool FAST>
__device__ __forceinline__ void pass_quad2(const CostArgs& a, const float4* __restrict__ srec, int count, int tp, int g, u64 MX, u64 MY, u64 MZ,
                                           double& acc0, double& acc1) {
    const float* __restrict__ I = a.info + 9 * (size_t)g;
    const float i0 = __ldg(I + 0), i1 = __ldg(I + 1), i2 = __ldg(I + 2), i3 = __ldg(I + 3), i4 = __ldg(I + 4), i5 = __ldg(I + 5), i6 = __ldg(I + 6),
                i7 = __ldg(I + 7), i8 = __ldg(I + 8);
    const float wk = __ldg(a.w + g);
    acc0 = acc1 = 0.0;
    int tprev = -1;
    u64 m[12];
#pragma unroll
    for (int c = 0; c < 12; ++c) m[c] = 0ull;
    const char* __restrict__ Mv = reinterpret_cast<const char*>(a.Mpair) + (size_t)tp * 96u;
    const unsigned rowbytes = (unsigned)a.Vp * 96u;
#define DMSA_QUAD_BODY2(jj)                                                                  \
    {                                                                                        \
        const float4 r = srec[(jj)];                                                         \
        const int t = __float_as_int(r.w);                                                   \
        u64 X, Y, Z;                                                                         \
        DMSA_XFORM2(r, X, Y, Z)                                                              \
        const u64 d0 = sub2(X, MX), d1 = sub2(Y, MY), d2 = sub2(Z, MZ);                      \
        const u64 t0 = mul2s(d0, wk), t1 = mul2s(d1, wk), t2 = mul2s(d2, wk);                \
        float r0a, r0b, r1a, r1b, r2a, r2b, sa, sb;                                          \
        DMSA_DOT3_2(mul2s(t0, i0), mul2s(t1, i3), mul2s(t2, i6), r0a, r0b)                   \
        DMSA_DOT3_2(mul2s(t0, i1), mul2s(t1, i4), mul2s(t2, i7), r1a, r1b)                   \
        DMSA_DOT3_2(mul2s(t0, i2), mul2s(t1, i5), mul2s(t2, i8), r2a, r2b)                   \
        DMSA_DOT3_2(mul2(pk2(r0a, r0b), d0), mul2(pk2(r1a, r1b), d1), mul2(pk2(r2a, r2b), d2), sa, sb) \
        acc0 += (double)sa;                                                                  \
        acc1 += (double)sb;                                                                  \
    }
    int j = 0;
    for (; j + 4 <= count; j += 4) {
        DMSA_QUAD_BODY2(j)
        DMSA_QUAD_BODY2(j + 1)
        DMSA_QUAD_BODY2(j + 2)
        DMSA_QUAD_BODY2(j + 3)
    }
    for (; j < count; ++j) DMSA_QUAD_BODY2(j)
#undef DMSA_QUAD_BODY2
}

// thread -> vector pair; threads beyond the last pair keep a valid table column and process no members
struct PairMap {
    int unit;  // set position / chunk index of this block
    int tp;
    bool has0, has1, fast;  // fast is block-uniform
};
__device__ __forceinline__ PairMap pair_map(const CostArgs& a) {
    PairMap pm;
    const int tid = threadIdx.x;
    pm.unit = a.split ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    pm.fast = a.split && (blockIdx.x & 1);
    pm.tp = pm.fast ? a.ns + tid : tid;
    const bool mine = pm.fast || !a.split || tid < a.ns;
    pm.has0 = mine && 2 * pm.tp < a.V;
    pm.has1 = mine && 2 * pm.tp + 1 < a.V;
    if (!pm.has0) pm.tp = pm.fast ? (a.V - 1) >> 1 : 0;
    return pm;
}

template <int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB) k_cost_fused2(CostArgs a) {
    DMSA_PDL_ENTER();
    __shared__ __align__(128) float4 srec[COST_CHUNK];
    __shared__ __align__(8) unsigned long long bar;
    const PairMap pm = pair_map(a);
    if (pm.unit >= total_sets(a.li)) return;
    const int g = a.order[pm.unit];
    const int kind = a.cell_kind[g];
    if (kind == 2) return;
    double* __restrict__ Eg = a.E + (size_t)g * a.Vld + 2 * pm.tp;
    if (kind == 0) {
        if (pm.has0) Eg[0] = 0.0;
        if (pm.has1) Eg[1] = 0.0;
        return;
    }
    const int n = a.cell_n[g];
    if (pm.fast)
        stage_partials(a, a.rec + a.cell_start[g], srec, n);
    else
        stage_records(srec, a.rec + a.cell_start[g], n, &bar);
    const int cnt = pm.has0 ? n : 0;
    PairSums S;
    if (pm.fast)
        pass_sum2<true>(a, srec, cnt, pm.tp, S);
    else
        pass_sum2<false>(a, srec, cnt, pm.tp, S);
    const float nf = (float)n;
    const u64 MX = pk2(fdiv_((float)S.x0, nf), fdiv_((float)S.x1, nf));  // DmsaOptimizer.h:254
    const u64 MY = pk2(fdiv_((float)S.y0, nf), fdiv_((float)S.y1, nf));
    const u64 MZ = pk2(fdiv_((float)S.z0, nf), fdiv_((float)S.z1, nf));
    double q0, q1;
    if (pm.fast)
        pass_quad2<true>(a, srec, cnt, pm.tp, g, MX, MY, MZ, q0, q1);
    else
        pass_quad2<false>(a, srec, cnt, pm.tp, g, MX, MY, MZ, q0, q1);
    if (pm.has0) Eg[0] = sqrt(fabs(q0));  // :267
    if (pm.has1) Eg[1] = sqrt(fabs(q1));
}

template <int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB) k_cost_sum2(CostArgs a) {
    DMSA_PDL_ENTER();
    __shared__ __align__(128) float4 srec[COST_CHUNK];
    __shared__ __align__(8) unsigned long long bar;
    __shared__ int s_last;
    const PairMap pm = pair_map(a);
    const int c = pm.unit;
    if (c >= *a.n_chunks) return;
    const Chunk ch = a.chunks[c];
    PairSums S;
    if (pm.fast) {
        stage_partials(a, a.rec + ch.start, srec, ch.count);
        pass_sum2<true>(a, srec, pm.has0 ? ch.count : 0, pm.tp, S);
    } else {
        stage_records(srec, a.rec + ch.start, ch.count, &bar);
        pass_sum2<false>(a, srec, pm.has0 ? ch.count : 0, pm.tp, S);
    }
    double* __restrict__ Sp = a.S_part + (size_t)c * 3 * a.Vld + 2 * pm.tp;
    if (pm.has0) {
        Sp[0] = S.x0;
        Sp[a.Vld] = S.y0;
        Sp[2 * (size_t)a.Vld] = S.z0;
    }
    if (pm.has1) {
        Sp[1] = S.x1;
        Sp[a.Vld + 1] = S.y1;
        Sp[2 * (size_t)a.Vld + 1] = S.z1;
    }
    const int g = ch.cell, nc = a.nchunk[g];
    if (nc <= MEAN_INLINE_MAX) return;
    int* const cnt1 = (pm.fast ? a.done1_f : a.done1) + g;  // the chunk blocks of one role reduce that role's vectors
    if (!last_block_of_set(cnt1, nc, &s_last)) return;
    if (pm.has0) {  // (an odd V: the second half of the last pair lands on a padding slot)
        const float nf = (float)a.cell_n[g];
        const size_t st3 = (size_t)3 * a.Vld;
        const double* __restrict__ Sq = a.S_part + (size_t)ch.first * st3 + 2 * pm.tp;
        float* __restrict__ mu = a.mu + (size_t)g * st3 + 2 * pm.tp;
        const double2 sx = reduce_chunks2(Sq, nc, st3), sy = reduce_chunks2(Sq + a.Vld, nc, st3), sz = reduce_chunks2(Sq + 2 * (size_t)a.Vld, nc, st3);
        mu[0] = fdiv_((float)sx.x, nf);
        mu[1] = fdiv_((float)sx.y, nf);
        mu[a.Vld] = fdiv_((float)sy.x, nf);
        mu[a.Vld + 1] = fdiv_((float)sy.y, nf);
        mu[2 * (size_t)a.Vld] = fdiv_((float)sz.x, nf);
        mu[2 * (size_t)a.Vld + 1] = fdiv_((float)sz.y, nf);
    }
    if (threadIdx.x == 0) *cnt1 = 0;
}

template <int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB) k_cost_quad2(CostArgs a) {
    DMSA_PDL_ENTER();
    __shared__ __align__(128) float4 srec[COST_CHUNK];
    __shared__ __align__(8) unsigned long long bar;
    __shared__ int s_last;
    const PairMap pm = pair_map(a);
    const int c = pm.unit;
    if (c >= *a.n_chunks) return;
    const Chunk ch = a.chunks[c];
    const int g = ch.cell;
    if (pm.fast)
        stage_partials(a, a.rec + ch.start, srec, ch.count);
    else
        stage_records(srec, a.rec + ch.start, ch.count, &bar);
    const int nc = a.nchunk[g], o = ch.first;
    const float nf = (float)a.cell_n[g];
    const size_t st3 = (size_t)3 * a.Vld;
    // (an odd V leaves the last thread's second half on a padding slot of S_part: computed, never stored)
    u64 MX, MY, MZ;
    if (nc <= MEAN_INLINE_MAX) {
        const double* __restrict__ Sp = a.S_part + (size_t)o * st3 + 2 * pm.tp;
        const double2 sx = reduce_chunks2(Sp, nc, st3), sy = reduce_chunks2(Sp + a.Vld, nc, st3), sz = reduce_chunks2(Sp + 2 * (size_t)a.Vld, nc, st3);
        MX = pk2(fdiv_((float)sx.x, nf), fdiv_((float)sx.y, nf));
        MY = pk2(fdiv_((float)sy.x, nf), fdiv_((float)sy.y, nf));
        MZ = pk2(fdiv_((float)sz.x, nf), fdiv_((float)sz.y, nf));
    } else {  // written by the last block of pass 1: adjacent floats of the pair, one 8-byte load per axis
        const u64* __restrict__ mu2 = reinterpret_cast<const u64*>(a.mu + (size_t)g * st3) + pm.tp;
        MX = mu2[0];
        MY = mu2[a.Vp];
        MZ = mu2[2 * (size_t)a.Vp];
    }
    double q0, q1;
    if (pm.fast)
        pass_quad2<true>(a, srec, pm.has0 ? ch.count : 0, pm.tp, g, MX, MY, MZ, q0, q1);
    else
        pass_quad2<false>(a, srec, pm.has0 ? ch.count : 0, pm.tp, g, MX, MY, MZ, q0, q1);
    double* __restrict__ Qc = a.Q + (size_t)c * a.Vld + 2 * pm.tp;
    if (pm.has0) Qc[0] = q0;
    if (pm.has1) Qc[1] = q1;
    int* const cnt2 = (pm.fast ? a.done_f : a.done) + g;
    if (!last_block_of_set(cnt2, nc, &s_last)) return;
    double* __restrict__ Eg = a.E + (size_t)g * a.Vld + 2 * pm.tp;
    const double2 qs = reduce_chunks2(a.Q + (size_t)o * a.Vld + 2 * pm.tp, nc, (size_t)a.Vld);
    if (pm.has0) Eg[0] = sqrt(fabs(qs.x));
    if (pm.has1) Eg[1] = sqrt(fabs(qs.y));
    if (threadIdx.x == 0) *cnt2 = 0;
}

// per-vector cost sum_r e[r][v]^2 (line search, DmsaOptimizer.h:171): COLSUM_PARTS row slices per vector, fixed reduction order
#define COLSUM_PARTS 16
__global__ void k_col_sumsq(const double* __restrict__ E, const LevelInfo* __restrict__ li, int n_extra, int Vld, double* __restrict__ part /*[9][COLSUM_PARTS]*/) {
    DMSA_PDL_ENTER();
    __shared__ double red[256];
    const int R = total_sets(li) + n_extra;
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
    DMSA_PDL_ENTER();
    const int v = threadIdx.x;
    if (v >= 9) return;
    double s = 0.0;
    for (int y = 0; y < COLSUM_PARTS; ++y) s += part[v * COLSUM_PARTS + y];
    out[v] = s;
}

// additional residual rows (IMU / gravity / odometry) behind the set rows: E[G + r][v] = extra[r][v]   (DmsaOptimizer.h:271-272)
__global__ void k_append_extra(double* __restrict__ E, const LevelInfo* __restrict__ li, const double* __restrict__ extra, int n_extra, int Vld) {
    DMSA_PDL_ENTER();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_extra * Vld) return;
    E[(size_t)total_sets(li) * Vld + i] = extra[i];
}

// Row partition of the J^T J kernels, computed on the device from the actual row count (the summation order, and with it
// every bit of H, depends on the partition: it must not depend on the host's grid-size guess).
#define JD_ROWS 32
#define JD_MAXBLK 148
__device__ __forceinline__ void jd_partition(int R, int& nblk, int& rpb) {
    nblk = max(1, min(JD_MAXBLK, (R + JD_ROWS - 1) / JD_ROWS));
    rpb = max(JD_ROWS, ((R + nblk - 1) / nblk + JD_ROWS - 1) / JD_ROWS * JD_ROWS);
    nblk = max(1, (R + rpb - 1) / rpb);
}
#define JTJ_T 32
#define JTJ_MAXSPLIT 64
__device__ __forceinline__ void jtj_partition(int R, int& nsplit, int& rps) {
    nsplit = max(1, min(JTJ_MAXSPLIT, (R + 127) / 128));
    rps = max(JTJ_T, ((R + nsplit - 1) / nsplit + JTJ_T - 1) / JTJ_T * JTJ_T);
    nsplit = max(1, (R + rps - 1) / rps);
}

// ---- [J e0]^T [J e0] : H = J^T J, g = J^T e0, err0 = e0^T e0 in one symmetric product ---------------------------
// Column c < P of the augmented matrix is the forward-difference column (E[:,c+1] - E[:,0]) / h (DmsaOptimizer.h:227),
// column P is e0.  Upper-triangular 32x32 tiles, split over row ranges; partials are reduced in fixed order.
__global__ void __launch_bounds__(256) k_jtj(const double* __restrict__ E, const LevelInfo* __restrict__ li, int n_extra, int Vld, int P, double inv_h,
                                             double* __restrict__ part /*[split][(P+1)*(P+1)]*/) {
    DMSA_PDL_ENTER();
    __shared__ double A[JTJ_T][JTJ_T + 1], B[JTJ_T][JTJ_T + 1];
    const int R = total_sets(li) + n_extra;
    int nsplit, rows_per_split;
    jtj_partition(R, nsplit, rows_per_split);
    if ((int)blockIdx.y >= nsplit) return;
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
__global__ void k_jtj_reduce(const double* __restrict__ part, const LevelInfo* __restrict__ li, int n_extra, int P, double* __restrict__ out) {
    DMSA_PDL_ENTER();
    const int n1 = P + 1;
    int nsplit, rps_;
    jtj_partition(total_sets(li) + n_extra, nsplit, rps_);
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

// ---- [J e0]^T [J e0] on the FP64 tensor cores (n1 = P + 1 <= 128) -----------------------------------------------
// The one genuine GEMM of the path: C = A^T A with A = [J | e0] (R x n1).  mma.sync.aligned.m8n8k4 (DMMA): the "A" operand
// is an 8 x 4 tile of A^T, i.e. A[k0 + t][8 ti + g], the "B" operand the 4 x 8 tile A[k0 + t][8 tj + g] (g = lane / 4,
// t = lane % 4): both are the same shared-memory access pattern.  A block owns a row range, stages 32-row panels of A in
// shared memory (row stride 132 doubles: conflict-free 8-byte fragment loads) and keeps ALL upper-triangular 8 x 8 output
// tiles in registers (<= 8 tiles per warp, 16 warps); per-block partials are reduced in fixed order by k_jtj_reduce8.
#define JD_LD 132
#define JD_T 512
#define JD_MAXN 128
__global__ void __launch_bounds__(JD_T, 1) k_jtj_dmma(const double* __restrict__ E, const LevelInfo* __restrict__ li, int n_extra, int Vld, int P, double inv_h,
                                                       double* __restrict__ part /*[block][n1*n1]*/) {
    DMSA_PDL_ENTER();
    __shared__ __align__(16) double As[JD_ROWS * JD_LD];
    __shared__ unsigned char tij[JD_MAXN / 8 * (JD_MAXN / 8 + 1) / 2][2];
    const int R = total_sets(li) + n_extra;
    int nblk_, rows_per_block;
    jd_partition(R, nblk_, rows_per_block);
    if ((int)blockIdx.x >= nblk_) return;
    const int n1 = P + 1, nt = (n1 + 7) >> 3, ntile = nt * (nt + 1) / 2;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
    for (int q = threadIdx.x; q < ntile; q += JD_T) {  // upper-triangular tile list, row by row
        int rem = q, ti = 0;
        while (rem >= nt - ti) {
            rem -= nt - ti;
            ++ti;
        }
        tij[q][0] = (unsigned char)ti;
        tij[q][1] = (unsigned char)(ti + rem);
    }
    double acc[8][2];
#pragma unroll
    for (int m = 0; m < 8; ++m) acc[m][0] = acc[m][1] = 0.0;
    const int r0 = blockIdx.x * rows_per_block, r1 = min(R, r0 + rows_per_block);
    for (int rb = r0; rb < r1; rb += JD_ROWS) {
        __syncthreads();
        for (int idx = threadIdx.x; idx < JD_ROWS * JD_MAXN; idx += JD_T) {
            const int rr = idx >> 7, cc = idx & 127, r = rb + rr;
            double v = 0.0;
            if (r < r1 && cc < n1) {
                const double e0 = E[(size_t)r * Vld];
                v = cc < P ? inv_h * (E[(size_t)r * Vld + cc + 1] - e0) : e0;  // DmsaOptimizer.h:227 | e0
            }
            As[rr * JD_LD + cc] = v;
        }
        __syncthreads();
#pragma unroll
        for (int m = 0; m < 8; ++m) {
            const int q = warp + (JD_T / 32) * m;
            if (q < ntile) {  // warp-uniform
                const double* __restrict__ pa = As + t * JD_LD + 8 * tij[q][0] + g;
                const double* __restrict__ pb = As + t * JD_LD + 8 * tij[q][1] + g;
#pragma unroll
                for (int k0 = 0; k0 < JD_ROWS; k0 += 4) {
                    const double a = pa[k0 * JD_LD], b = pb[k0 * JD_LD];
                    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                                 : "+d"(acc[m][0]), "+d"(acc[m][1])
                                 : "d"(a), "d"(b));
                }
            }
        }
    }
    double* __restrict__ out = part + (size_t)blockIdx.x * n1 * n1;
#pragma unroll
    for (int m = 0; m < 8; ++m) {
        const int q = warp + (JD_T / 32) * m;
        if (q < ntile) {
            const int i = 8 * tij[q][0] + g, j = 8 * tij[q][1] + 2 * t;
            if (i < n1 && j < n1) out[(size_t)i * n1 + j] = acc[m][0];
            if (i < n1 && j + 1 < n1) out[(size_t)i * n1 + j + 1] = acc[m][1];
        }
    }
}
// out = [H (P*P row-major) | g (P) | err0] from the per-block partials of k_jtj_dmma (8 x 8 tiles, lower tiles mirrored).
// blockDim = (32, 8): 32 consecutive output elements x 8 interleaved partial streams (partial k goes to stream k % 8,
// ascending k), the streams are combined in ascending order: fixed summation order, coalesced 256-byte reads.
__global__ void __launch_bounds__(256) k_jtj_reduce8(const double* __restrict__ part, const LevelInfo* __restrict__ li, int n_extra, int P,
                                                     double* __restrict__ out) {
    DMSA_PDL_ENTER();
    __shared__ double red[8][33];
    const int n1 = P + 1;
    int nsplit, rpb_;
    jd_partition(total_sets(li) + n_extra, nsplit, rpb_);
    const int q = blockIdx.x * 32 + threadIdx.x;
    const bool ok = q < n1 * n1;
    const int i = ok ? q / n1 : 0, j = ok ? q % n1 : 0;
    int ii = i, jj = j;
    if ((i >> 3) > (j >> 3)) {
        ii = j;
        jj = i;
    }
    double s = 0.0;
    if (ok) {
        const double* __restrict__ src = part + (size_t)ii * n1 + jj;
#pragma unroll 4
        for (int k = threadIdx.y; k < nsplit; k += 8) s += src[(size_t)k * n1 * n1];
    }
    red[threadIdx.y][threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.y != 0 || !ok) return;
    double tsum = 0.0;
#pragma unroll
    for (int y = 0; y < 8; ++y) tsum += red[y][threadIdx.x];
    if (i < P && j < P)
        out[(size_t)i * P + j] = tsum;
    else if (i < P && j == P)
        out[(size_t)P * P + i] = tsum;
    else if (i == P && j == P)
        out[(size_t)P * P + P] = tsum;
}

// ---- base-pose world points (updateGlobalPoints): scan points through column v of the table ----------------------
// ContinuousTrajectory.h:137-155 | MapManagement.h:140-147
__global__ void k_transform_points(const float4* __restrict__ local, const int* __restrict__ tid, int n, const float4* __restrict__ Mtab, int Vld,
                                   int v, float4* __restrict__ world, const float4* __restrict__ normal_l, float4* __restrict__ normal_w) {
    DMSA_PDL_ENTER();
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
