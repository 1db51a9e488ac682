// kernels_sets.cuh — device-side Gaussian-set construction.
//
// Restates on the GPU:
//   DmsaOptimizer.h:275-350  createGaussianSets  (voxel partition, acceptance test n >= minPts && ids not all equal)
//   PCL 1.10 OctreePointCloud::addPointsFromInputCloud / adoptBoundingBoxToPoint / getKeyBitSize /
//            genOctreeKeyforPoint and the depth-first leaf iterator (third-party, restated from the published algorithm)
//   Gaussians.h:130-168      addPointSet (covariance), :181-201 limitCovariance, :170-179 updateRebalancingWeights
//
// Pipeline (both resolution levels per launch): anchor -> voxel keys (+ per-block key boxes) -> octree root growth replay ->
// [kernels_sort.cuh: Morton codes (x most significant == PCL child index (x<<2)|(y<<1)|z) -> stable radix sort (members stay
// in ascending point index) -> run heads / ring test -> accept / number / emit -> gather member records] -> per-set
// covariance / eigen clamp / information matrix.  All of it streams the point set a constant number of times.
#pragma once
#include <cuda_runtime.h>

#include "pdl.cuh"

#include <cstdint>

namespace dmsa {

#define DMSA_KEYS_BLOCK 256

#ifdef DMSA_ALLOW_FMA
// EXPERIMENT ONLY (scripts/fma_deviation.py builds a second library with -fmad=true): plain operators, so that the compiler
// contracts a * b + c into FFMA — the measurement of what fusing would do to H and g (profiles/r02_fma_deviation.json).
__device__ __forceinline__ float fmul_(float a, float b) { return a * b; }
__device__ __forceinline__ float fadd_(float a, float b) { return a + b; }
__device__ __forceinline__ float fsub_(float a, float b) { return a - b; }
#else
__device__ __forceinline__ float fmul_(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fadd_(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub_(float a, float b) { return __fsub_rn(a, b); }
#endif
__device__ __forceinline__ float fdiv_(float a, float b) { return __fdiv_rn(a, b); }

struct LevelInfo {
    double min0[3];   // lattice anchor: octree min after the first point (PCL getKeyBitSize on the empty tree)
    double max0[3];
    double res;
    int first;        // index of the first finite point
    int kmin[3], kmax[3];
    long long lo[3];  // octree root origin in anchor-relative key units after all growth steps
    int depth;
    int n_valid;      // finite points
    int R;            // raw (non-empty) leaves
    int G;            // accepted sets of this level
    int gbase;        // first set index of this level in the store
    int error;        // 1: depth > 21
    int bound;        // most sets the consumers of this build may touch (grids / row buffers were sized for it)
};

// number of accepted sets of both levels (a level that is switched off keeps the zeroed record of the per-build memset):
// the kernels behind the set build read it from the device so that the host does not have to wait for it
// (clamped to the bound the host sized its buffers for: a deferred build whose guess was too small must stay in bounds;
// the host sees the unclamped counts in the read-back and redoes the iteration)
__device__ __forceinline__ int total_sets(const LevelInfo* __restrict__ li) { return min(li[0].G + li[1].G, max(li[0].bound, li[1].bound)); }

// ---- anchor (one thread) ---------------------------------------------------------------------
struct LevelPlan {  // the resolution levels built in this pass (DmsaOptimizer.h:81-86: a factor <= FLT_MIN disables a level)
    int n;
    int level[2];
    float res[2];
};
__global__ void k_anchor(const float4* __restrict__ world, int N, LevelPlan plan, LevelInfo* infos, int bound) {
    DMSA_PDL_ENTER();
    if (blockIdx.x != 0) return;
    {  // both level records start from zero (a disabled level keeps G = R = 0); launched with one warp
        unsigned int* w = reinterpret_cast<unsigned int*>(infos);
        for (unsigned int i = threadIdx.x; i < 2 * sizeof(LevelInfo) / sizeof(unsigned int); i += blockDim.x) w[i] = 0u;
        __syncwarp();
    }
    if ((int)threadIdx.x >= plan.n) return;
    LevelInfo* info = infos + plan.level[threadIdx.x];
    info->bound = bound;
    const double res = (double)plan.res[threadIdx.x];  // OctreePointCloud(const double resolution)
    const double minValue = 1.1920928955078125e-07;  // std::numeric_limits<float>::epsilon()
    int i = 0;
    while (i < N) {
        float4 p = world[i];
        if (isfinite(p.x) && isfinite(p.y) && isfinite(p.z)) break;
        ++i;
    }
    info->first = i;
    info->res = res;
    info->error = 0;
    info->R = 0;
    info->G = 0;
    info->n_valid = 0;
    for (int a = 0; a < 3; ++a) {
        info->kmin[a] = 2147483647;
        info->kmax[a] = -2147483647 - 1;
        info->lo[a] = 0;
    }
    info->depth = 1;
    if (i >= N) {
        for (int a = 0; a < 3; ++a) info->min0[a] = info->max0[a] = 0.0;
        return;
    }
    float4 p = world[i];
    double c[3] = {(double)p.x, (double)p.y, (double)p.z};
    // adoptBoundingBoxToPoint (box undefined): min = p - res/2, max = p + res/2; getKeyBitSize(): depth 1,
    // oversize = (2 res - (max - min)) / 2 applied on both sides when > eps
    const double side = 2.0 * res;
    for (int a = 0; a < 3; ++a) {
        double mn = c[a] - res / 2, mx = c[a] + res / 2;
        double over = (side - (mx - mn)) / 2.0;
        if (over > minValue) {
            mn -= over;
            mx += over;
        }
        info->min0[a] = mn;
        info->max0[a] = mx;
    }
}

// ---- voxel keys: key_a = floor((double(x_a) - min0_a) / res)   (genOctreeKeyforPoint, root-growth invariant) -------
// Per 256-point block it also records the key box (bbmin/bbmax) and the key box of "edge" points (ebmin/ebmax):
// points closer to a voxel face than PCL's bounding-box fudge (float eps on the upper side) or than the rounding
// noise of the lattice arithmetic; only those can make the exact double test of k_root disagree with the integer test.
// grid = (blocks of 256 points, levels): both resolution levels in one launch.
// The launch also zeroes up to two word ranges the NEXT phase needs cleared (the control block of the sort / scans and the
// ring-test flags: ~7 MB at BASELINE config 2) — spread over all blocks it costs nothing here and takes two memsets off the
// dependent chain between k_root and k_sort_prepare.
struct ZeroRanges {
    unsigned int* p[2];
    size_t words[2];
};
__global__ void k_keys(const float4* __restrict__ world, int N, LevelPlan plan, LevelInfo* __restrict__ infos, int* __restrict__ keys_all,
                       int* __restrict__ bb_all, int nb, ZeroRanges zr) {
    DMSA_PDL_ENTER();
    {
        const size_t nthreads = (size_t)gridDim.x * gridDim.y * blockDim.x;
        const size_t t0 = ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * blockDim.x + threadIdx.x;
#pragma unroll
        for (int r = 0; r < 2; ++r)
            for (size_t w = t0; w < zr.words[r]; w += nthreads) zr.p[r][w] = 0u;
    }
    const int lvl = plan.level[blockIdx.y];
    LevelInfo* __restrict__ info = infos + lvl;
    int* __restrict__ keys = keys_all + (size_t)3 * N * lvl;
    int* __restrict__ bbmin = bb_all + (size_t)12 * nb * lvl;
    int* __restrict__ bbmax = bbmin + 3 * nb;
    int* __restrict__ ebmin = bbmin + 6 * nb;
    int* __restrict__ ebmax = bbmin + 9 * nb;
    __shared__ int smin[3][DMSA_KEYS_BLOCK / 32], smax[3][DMSA_KEYS_BLOCK / 32];
    __shared__ int semin[3][DMSA_KEYS_BLOCK / 32], semax[3][DMSA_KEYS_BLOCK / 32];
    int ek[3] = {2147483647, 2147483647, 2147483647};
    int ekx[3] = {-2147483647 - 1, -2147483647 - 1, -2147483647 - 1};
    const int i = blockIdx.x * DMSA_KEYS_BLOCK + threadIdx.x;
    const double res = info->res;
    int k[3] = {2147483647, 2147483647, 2147483647};
    int kx[3] = {-2147483647 - 1, -2147483647 - 1, -2147483647 - 1};
    bool valid = false;
    if (i < N) {
        float4 p = world[i];
        valid = isfinite(p.x) && isfinite(p.y) && isfinite(p.z);
        if (valid) {
            double c[3] = {(double)p.x, (double)p.y, (double)p.z};
            const double inv_res = 1.0 / res;
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                // PCL divides: key = floor((x - min) / res).  The product with 1 / res differs from the quotient by a few ulp
                // (< 1e-15 |q|), so away from the integers both floor to the same key; only within `guard` of a voxel face
                // the exact division decides (rare), and the edge-point test below always sees the exact fraction there.
                const double d = c[a] - info->min0[a];
                double q = d * inv_res;
                double fl = floor(q);
                double frac = q - fl;
                const double guard = 4e-9 + 4e-15 * fabs(q);
                if (frac < guard || frac > 1.0 - guard) {
                    q = d / res;
                    fl = floor(q);
                    frac = q - fl;
                }
                k[a] = kx[a] = (int)fl;
                if (frac < 1e-9 || (1.0 - frac) * res <= 2.4e-7) ek[a] = ekx[a] = (int)fl;
            }
        }
        keys[3 * (size_t)i + 0] = valid ? k[0] : (-2147483647 - 1);
        keys[3 * (size_t)i + 1] = valid ? k[1] : (-2147483647 - 1);
        keys[3 * (size_t)i + 2] = valid ? k[2] : (-2147483647 - 1);
    }
    // block bounding box of keys
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        int mn = k[a], mx = kx[a];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
            mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        }
        int emn = ek[a], emx = ekx[a];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            emn = min(emn, __shfl_xor_sync(0xffffffffu, emn, o));
            emx = max(emx, __shfl_xor_sync(0xffffffffu, emx, o));
        }
        if ((threadIdx.x & 31) == 0) {
            smin[a][threadIdx.x >> 5] = mn;
            smax[a][threadIdx.x >> 5] = mx;
            semin[a][threadIdx.x >> 5] = emn;
            semax[a][threadIdx.x >> 5] = emx;
        }
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        int a = threadIdx.x;
        int mn = smin[a][0], mx = smax[a][0], emn = semin[a][0], emx = semax[a][0];
        for (int w = 1; w < DMSA_KEYS_BLOCK / 32; ++w) {
            mn = min(mn, smin[a][w]);
            mx = max(mx, smax[a][w]);
            emn = min(emn, semin[a][w]);
            emx = max(emx, semax[a][w]);
        }
        bbmin[3 * blockIdx.x + a] = mn;
        bbmax[3 * blockIdx.x + a] = mx;
        ebmin[3 * blockIdx.x + a] = emn;
        ebmax[3 * blockIdx.x + a] = emx;
    }
}

// ---- octree root growth replay (adoptBoundingBoxToPoint for every later point, in index order) -----------------
// One block.  The box only grows, so violators are found in increasing index order; per-block key boxes skip
// blocks that cannot contain a violator (conservative integer test), candidates are re-tested exactly in double.
__global__ void k_root(const float4* __restrict__ world, int N, LevelPlan plan, LevelInfo* __restrict__ infos, const int* __restrict__ bb_all,
                       int nb_) {
    DMSA_PDL_ENTER();
    const int lvl = plan.level[blockIdx.x];  // one block per level
    LevelInfo* __restrict__ info = infos + lvl;
    const int* __restrict__ bbmin = bb_all + (size_t)12 * nb_ * lvl;
    const int* __restrict__ bbmax = bbmin + 3 * nb_;
    const int* __restrict__ ebmin = bbmin + 6 * nb_;
    const int* __restrict__ ebmax = bbmin + 9 * nb_;
    __shared__ double mn[3], mx[3];
    __shared__ long long lo[3];
    __shared__ int depth, found, cand;
    const double minValue = 1.1920928955078125e-07;
    const int nb = (N + DMSA_KEYS_BLOCK - 1) / DMSA_KEYS_BLOCK;
    if (threadIdx.x == 0) {
        for (int a = 0; a < 3; ++a) {
            mn[a] = info->min0[a];
            mx[a] = info->max0[a];
            lo[a] = 0;
        }
        depth = 1;
    }
    __syncthreads();
    const double res = info->res;
    int b = 0;  // first block not yet cleared
    while (b < nb) {
        // 1. first candidate block >= b
        if (threadIdx.x == 0) cand = nb;
        __syncthreads();
        for (int base = b; base < nb; base += blockDim.x) {
            int bi = base + threadIdx.x;
            bool c = false;
            if (bi < nb) {
                long long side = ((long long)1 << depth);
#pragma unroll
                for (int a = 0; a < 3; ++a) {
                    int kmn = bbmin[3 * bi + a], kmx = bbmax[3 * bi + a];
                    if (kmn <= kmx) c = c || ((long long)kmn < lo[a]) || ((long long)kmx >= lo[a] + side);
                    int emn = ebmin[3 * bi + a], emx = ebmax[3 * bi + a];
                    if (emn <= emx) c = c || ((long long)emn <= lo[a]) || ((long long)emx >= lo[a] + side - 1);
                }
            }
            if (c) atomicMin(&cand, bi);
            __syncthreads();
            const int cv = *reinterpret_cast<volatile int*>(&cand);
            __syncthreads();
            if (cv < nb) break;
        }
        __syncthreads();
        const int cnd = *reinterpret_cast<volatile int*>(&cand);
        if (cnd >= nb) break;
        b = cnd;
        // 2. exact test of the candidate block's points, repeatedly (each growth step may leave later violators)
        while (true) {
            if (threadIdx.x == 0) found = 2147483647;
            __syncthreads();
            for (int off = threadIdx.x; off < DMSA_KEYS_BLOCK; off += blockDim.x) {
                int i = b * DMSA_KEYS_BLOCK + off;
                if (i < N) {
                    float4 p = world[i];
                    if (isfinite(p.x) && isfinite(p.y) && isfinite(p.z)) {
                        double c[3] = {(double)p.x, (double)p.y, (double)p.z};
                        bool viol = false;
#pragma unroll
                        for (int a = 0; a < 3; ++a) viol = viol || (c[a] < mn[a]) || (c[a] >= mx[a]);
                        if (viol) atomicMin(&found, i);
                    }
                }
            }
            __syncthreads();
            // (volatile: keeps the compiler from fetching the neighbouring `depth` word in the same wide shared load, which
            //  thread 0 rewrites below before the next barrier)
            const int fnd = *reinterpret_cast<volatile int*>(&found);
            if (fnd == 2147483647) break;
            if (threadIdx.x == 0) {
                float4 p = world[fnd];
                double c[3] = {(double)p.x, (double)p.y, (double)p.z};
                while (true) {
                    bool up[3], any = false;
                    for (int a = 0; a < 3; ++a) {
                        bool lowv = c[a] < mn[a];
                        up[a] = c[a] >= mx[a];
                        any = any || lowv || up[a];
                    }
                    if (!any) break;
                    double side = (double)(1u << depth) * res;
                    for (int a = 0; a < 3; ++a)
                        if (!up[a]) {
                            mn[a] -= side;
                            lo[a] -= (long long)1 << depth;
                        }
                    depth++;
                    side = (double)(1u << depth) * res - minValue;
                    for (int a = 0; a < 3; ++a) mx[a] = mn[a] + side;
                    if (depth > 21) break;
                }
            }
            __syncthreads();
            if (depth > 21) break;
        }
        if (depth > 21) break;
        b = b + 1;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int a = 0; a < 3; ++a) info->lo[a] = lo[a];
        info->depth = depth;
        info->error = depth > 21 ? 1 : 0;
    }
}

__device__ __forceinline__ unsigned long long spread3(unsigned long long x) {  // 21 bits -> every third bit
    x &= 0x1fffffull;
    x = (x | (x << 32)) & 0x1f00000000ffffull;
    x = (x | (x << 16)) & 0x1f0000ff0000ffull;
    x = (x | (x << 8)) & 0x100f00f00f00f00full;
    x = (x | (x << 4)) & 0x10c30c30c30c30c3ull;
    x = (x | (x << 2)) & 0x1249249249249249ull;
    return x;
}

// Acceptance of a leaf (DmsaOptimizer.h:307) and its default emission plan: one unsplit set.
// out_cnt[c] in {0,1,2} sets are emitted for leaf c; sub_* describe them (2 slots per leaf).
__global__ void k_accept(const int* __restrict__ raw_start, const int* __restrict__ raw_diff, const LevelInfo* __restrict__ info, int minPts,
                         int* __restrict__ acc_flag, int* __restrict__ out_cnt, int* __restrict__ sub_start, int* __restrict__ sub_n,
                         int* __restrict__ sub_code) {
    DMSA_PDL_ENTER();
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= info->R) return;
    const int s = raw_start[c];
    const int n = raw_start[c + 1] - s;
    const int acc = (n >= minPts && raw_diff[c]) ? 1 : 0;
    acc_flag[c] = acc;
    out_cnt[c] = acc;
    sub_start[2 * c] = s;
    sub_n[2 * c] = n;
    sub_code[2 * c] = 0;
}

// ---- splitSet<PointCloud<PointNormal>> (Gaussians.h:27-85; keyframe pass with gauss_split) --------------------------------
// Pair search: over ordered pairs (a != b) of a leaf's members the minimum of ||n_a + n_b|| (float, first minimum in loop
// order wins == lexicographic minimum of (value, position a, position b)).  256 first members per block x all second members.
struct SplitTile {
    int cell, i1_start;
};
// Pre-filter of the O(n^2) pair search (exact, not a heuristic): for ANY vector r with |r| <= 1,
//     ||n_a + n_b|| >= (n_a + n_b) . r >= 2 min_j (n_j . r),
// so a leaf whose normals all satisfy n_j . r > 0.3 has min ||n_a + n_b|| > 0.6 > 0.5 and splitSet leaves it unsplit
// (Gaussians.h:54) whatever the arg-min pair is.  r = the centre of the leaf's normal bounding box, normalised (rounded
// down in length): independent of how many points each surface contributes, so walls meeting at a corner pass.  Only the
// leaves that fail (surfaces seen from both sides, the ones that really split) run the pair search.
__device__ __forceinline__ int f2ord(float f) {  // order-preserving float -> int
    const int i = __float_as_int(f);
    return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float ord2f(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }
__global__ void k_split_nbox(const int* __restrict__ idx, const int* __restrict__ scan, const int* __restrict__ acc_flag, const LevelInfo* __restrict__ info,
                             const float4* __restrict__ normal_w, int* __restrict__ nbox /*[R][6]: ordered-int min xyz, max xyz*/) {
    DMSA_PDL_ENTER();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const bool in = i < info->n_valid;
    const int c = in ? scan[i] - 1 : -1 - lane;  // (distinct negative ids: never merged)
    const bool act = in && acc_flag[c];
    int mn[3] = {2147483647, 2147483647, 2147483647}, mx[3] = {-2147483647 - 1, -2147483647 - 1, -2147483647 - 1};
    if (act) {
        const float4 n = normal_w[idx[i]];
        mn[0] = mx[0] = f2ord(n.x);
        mn[1] = mx[1] = f2ord(n.y);
        mn[2] = mx[2] = f2ord(n.z);
    }
    // segmented min / max over the contiguous runs of equal leaf id inside the warp
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int c2 = __shfl_down_sync(0xffffffffu, c, o);
        const bool take = lane + o < 32 && c2 == c;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const int m2 = __shfl_down_sync(0xffffffffu, mn[a], o), x2 = __shfl_down_sync(0xffffffffu, mx[a], o);
            if (take) {
                mn[a] = min(mn[a], m2);
                mx[a] = max(mx[a], x2);
            }
        }
    }
    const int cprev = __shfl_up_sync(0xffffffffu, c, 1);
    if (act && (lane == 0 || cprev != c)) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            atomicMin(nbox + 6 * (size_t)c + a, mn[a]);
            atomicMax(nbox + 6 * (size_t)c + 3 + a, mx[a]);
        }
    }
}
__global__ void k_split_prefilter(const int* __restrict__ idx, const int* __restrict__ scan, const int* __restrict__ acc_flag, const LevelInfo* __restrict__ info,
                                  const float4* __restrict__ normal_w, const int* __restrict__ nbox, int* __restrict__ search /*[R]: 1 = run the pair search*/) {
    DMSA_PDL_ENTER();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= info->n_valid) return;
    const int c = scan[i] - 1;
    if (!acc_flag[c]) return;
    const int* __restrict__ b = nbox + 6 * (size_t)c;
    const float sx = 0.5f * (ord2f(b[0]) + ord2f(b[3])), sy = 0.5f * (ord2f(b[1]) + ord2f(b[4])), sz = 0.5f * (ord2f(b[2]) + ord2f(b[5]));
    const float len = sqrtf(sx * sx + sy * sy + sz * sz);
    bool ok = false;
    if (len > 1e-3f && len < 3.0e38f) {
        const float inv = 1.0f / (len * 1.00001f);  // |r| < 1
        const float4 n = normal_w[idx[i]];
        const float d = (n.x * sx + n.y * sy + n.z * sz) * inv;
        ok = d > 0.3f;  // (NaN compares false -> searched)
    }
    if (!ok) search[c] = 1;
}
// nbox initial values: min slots = INT_MAX, max slots = INT_MIN
__global__ void k_split_nbox_init(int* __restrict__ nbox, int* __restrict__ search, int n_cells, const LevelInfo* __restrict__ info) {
    DMSA_PDL_ENTER();
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_cells || c > info->R) return;  // (only the leaves of this build)
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        nbox[6 * (size_t)c + a] = 2147483647;
        nbox[6 * (size_t)c + 3 + a] = -2147483647 - 1;
    }
    search[c] = 0;
}
__global__ void k_split_tile_counts(const int* __restrict__ raw_start, const int* __restrict__ acc_flag, const int* __restrict__ search,
                                    const LevelInfo* __restrict__ info, int* __restrict__ ntile) {
    DMSA_PDL_ENTER();
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int R = info->R;
    if (c > R) return;
    ntile[c] = (c < R && acc_flag[c] && search[c]) ? (raw_start[c + 1] - raw_start[c] + 255) / 256 : 0;
}
__global__ void k_split_tile_fill(const int* __restrict__ ntile, const int* __restrict__ tile_off, const LevelInfo* __restrict__ info,
                                  SplitTile* __restrict__ tiles) {
    DMSA_PDL_ENTER();
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= info->R) return;
    const int nt = ntile[c], o = tile_off[c];
    for (int t = 0; t < nt; ++t) {
        SplitTile st;
        st.cell = c;
        st.i1_start = t * 256;
        tiles[o + t] = st;
    }
}
__device__ __forceinline__ bool pair_less(float v, int i, int j, float v2, int i2, int j2) {
    return v < v2 || (v == v2 && (i < i2 || (i == i2 && j < j2)));
}
__global__ void __launch_bounds__(256) k_split_pairs(const SplitTile* __restrict__ tiles, const int* __restrict__ tile_off, const LevelInfo* __restrict__ info,
                                                     const int* __restrict__ raw_start, const int* __restrict__ sidx,
                                                     const float4* __restrict__ normal_w, float* __restrict__ best_v, int* __restrict__ best_i,
                                                     int* __restrict__ best_j) {
    DMSA_PDL_ENTER();
    __shared__ float sx[256], sy[256], sz[256];
    __shared__ float rv[256];
    __shared__ int ri[256], rj[256];
    const int total = tile_off[info->R];
  for (int t = blockIdx.x; t < total; t += gridDim.x) {  // persistent grid: the tile count lives on the device (block-uniform loop)
    __syncthreads();
    const SplitTile tl = tiles[t];
    const int s = raw_start[tl.cell], n = raw_start[tl.cell + 1] - s;
    const int i1 = tl.i1_start + threadIdx.x;
    float ax = 0.f, ay = 0.f, az = 0.f;
    if (i1 < n) {
        const float4 a = normal_w[sidx[s + i1]];
        ax = a.x;
        ay = a.y;
        az = a.z;
    }
    // Two exact reductions of the reference's double loop over ordered pairs (Gaussians.h:33-52):
    //  * ||n_a + n_b|| is bitwise symmetric in (a, b) (float addition commutes) and the first strict minimum in loop order is
    //    the lexicographically smallest pair, so only the pairs a < b are visited;
    //  * sqrt is monotone: a candidate whose squared length is not below the best one's cannot have a smaller norm, so the
    //    square root is taken (and compared, strictly, like the reference's `<`) only for the few candidates that are.
    float bv = 3.402823466e+38f;  // std::numeric_limits<float>::max(), Gaussians.h:31
    float bs = __int_as_float(0x7f800000);
    int bj = 0x7fffffff;
    for (int j0 = tl.i1_start; j0 < n; j0 += 256) {
        const int jj = j0 + threadIdx.x;
        if (jj < n) {
            const float4 b = normal_w[sidx[s + jj]];
            sx[threadIdx.x] = b.x;
            sy[threadIdx.x] = b.y;
            sz[threadIdx.x] = b.z;
        }
        __syncthreads();
        const int lim = min(256, n - j0);
        if (i1 < n) {
            for (int q = (j0 == tl.i1_start ? (int)threadIdx.x + 1 : 0); q < lim; ++q) {
                const float ux = fadd_(ax, sx[q]), uy = fadd_(ay, sy[q]), uz = fadd_(az, sz[q]);
                const float ss = fadd_(fmul_(ux, ux), fadd_(fmul_(uy, uy), fmul_(uz, uz)));  // (n1 + n2).squaredNorm()
                if (ss < bs) {
                    const float v = __fsqrt_rn(ss);  // .norm()
                    if (v < bv) {
                        bv = v;
                        bs = ss;
                        bj = j0 + q;
                    }
                }
            }
        }
        __syncthreads();
    }
    rv[threadIdx.x] = bv;
    ri[threadIdx.x] = (i1 < n) ? i1 : 0x7fffffff;
    rj[threadIdx.x] = bj;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
            if (pair_less(rv[threadIdx.x + o], ri[threadIdx.x + o], rj[threadIdx.x + o], rv[threadIdx.x], ri[threadIdx.x], rj[threadIdx.x])) {
                rv[threadIdx.x] = rv[threadIdx.x + o];
                ri[threadIdx.x] = ri[threadIdx.x + o];
                rj[threadIdx.x] = rj[threadIdx.x + o];
            }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        best_v[t] = rv[0];
        best_i[t] = ri[0];
        best_j[t] = rj[0];
    }
  }
}
// One block per accepted leaf: reduce the tiles' winners; if min <= 0.5 split the members by the nearer reference normal
// (stable partition of the leaf's range of sidx through `scratch`), ring test of the first half, the reference's two
// acceptance tests (DmsaOptimizer.h:319 and the :327-331 quirk: the second half re-tests the FIRST half's rings).
__global__ void __launch_bounds__(256) k_split_decide(const int* __restrict__ acc_flag, const int* __restrict__ ntile, const int* __restrict__ tile_off,
                                                      const LevelInfo* __restrict__ info, const int* __restrict__ raw_start, int* __restrict__ sidx,
                                                      int* __restrict__ scratch, const float4* __restrict__ normal_w, const int* __restrict__ ring,
                                                      const float* __restrict__ best_v, const int* __restrict__ best_i, const int* __restrict__ best_j,
                                                      int minPts, int* __restrict__ out_cnt, int* __restrict__ sub_start, int* __restrict__ sub_n,
                                                      int* __restrict__ sub_code) {
    DMSA_PDL_ENTER();
    __shared__ int scan[256];
    __shared__ int s_rmin, s_rmax;
    __shared__ float s_bv;
    __shared__ int s_bi, s_bj;
  for (int c = blockIdx.x; c < info->R; c += gridDim.x) {  // block-uniform loop over leaves
    __syncthreads();
    if (!acc_flag[c]) continue;
    const int s = raw_start[c], n = raw_start[c + 1] - s;
    if (threadIdx.x == 0) {
        float bv = 3.402823466e+38f;
        int bi = 0x7fffffff, bj = 0x7fffffff;
        const int o = tile_off[c], nt = ntile[c];
        for (int t = 0; t < nt; ++t)
            if (pair_less(best_v[o + t], best_i[o + t], best_j[o + t], bv, bi, bj)) {
                bv = best_v[o + t];
                bi = best_i[o + t];
                bj = best_j[o + t];
            }
        s_bv = bv;
        s_bi = bi;
        s_bj = bj;
        s_rmin = 2147483647;
        s_rmax = -2147483647 - 1;
    }
    __syncthreads();
    if (s_bv > 0.5f) continue;  // Gaussians.h:54: no opposite normals -> the default plan (one unsplit set) stands
    const float4 ra = normal_w[sidx[s + s_bi]], rb = normal_w[sidx[s + s_bj]];
    __syncthreads();
    // stable partition: first-half members to the front of scratch in order, second-half members (reversed) to the back
    int base1 = 0;
    for (int j0 = 0; j0 < n; j0 += 256) {
        const int j = j0 + threadIdx.x;
        int f = 0, id = 0;
        if (j < n) {
            id = sidx[s + j];
            const float4 m = normal_w[id];
            const float ax = fsub_(ra.x, m.x), ay = fsub_(ra.y, m.y), az = fsub_(ra.z, m.z);
            const float bx = fsub_(rb.x, m.x), by = fsub_(rb.y, m.y), bz = fsub_(rb.z, m.z);
            const float d1 = __fsqrt_rn(fadd_(fmul_(ax, ax), fadd_(fmul_(ay, ay), fmul_(az, az))));
            const float d2 = __fsqrt_rn(fadd_(fmul_(bx, bx), fadd_(fmul_(by, by), fmul_(bz, bz))));
            f = d1 < d2 ? 1 : 0;  // Gaussians.h:75
        }
        scan[threadIdx.x] = f;
        __syncthreads();
        for (int o = 1; o < 256; o <<= 1) {  // inclusive Hillis-Steele scan
            int v = threadIdx.x >= o ? scan[threadIdx.x - o] : 0;
            __syncthreads();
            scan[threadIdx.x] += v;
            __syncthreads();
        }
        const int incl = scan[threadIdx.x], tot = scan[255];
        if (j < n) {
            if (f) {
                scratch[s + base1 + incl - 1] = id;
                atomicMin(&s_rmin, ring[id]);
                atomicMax(&s_rmax, ring[id]);
            } else {
                const int k2 = j - (base1 + incl);  // rank among the second-half members
                scratch[s + n - 1 - k2] = id;
            }
        }
        base1 += tot;
        __syncthreads();
    }
    const int n1 = base1, n2 = n - n1;
    for (int j = threadIdx.x; j < n1; j += 256) sidx[s + j] = scratch[s + j];
    for (int j = threadIdx.x; j < n2; j += 256) sidx[s + n1 + j] = scratch[s + n - 1 - j];
    if (threadIdx.x == 0) {
        const bool ring1 = s_rmax != s_rmin;
        const int ok1 = (n1 > minPts && ring1) ? 1 : 0;
        const int ok2 = (n2 > minPts && ring1) ? 1 : 0;  // quirk: the FIRST half's ring test again
        out_cnt[c] = ok1 + ok2;
        int e = 0;
        if (ok1) {
            sub_start[2 * c + e] = s;
            sub_n[2 * c + e] = n1;
            sub_code[2 * c + e] = 1;
            ++e;
        }
        if (ok2) {
            sub_start[2 * c + e] = s + n1;
            sub_n[2 * c + e] = n2;
            sub_code[2 * c + e] = 2;
        }
    }
  }
}

struct CellStore {
    int* start;   // first member record (index into the 2N-long sorted member arrays)
    int* n;       // members
    int* level;
    int* key;     // 3 per set
    int* sub;
    float* info;  // 9 per set, row-major
    float* w0;    // (1/n) * observation weight
    float* w;     // rebalancing weight
};

// cyclic Jacobi, symmetric 3x3, double
__device__ inline void jacobi3(const double Ain[9], double l[3], double V[9]) {
    double A[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) {
        A[i] = Ain[i];
        V[i] = (i % 4 == 0) ? 1.0 : 0.0;
    }
    for (int sweep = 0; sweep < 12; ++sweep) {
        double off = A[1] * A[1] + A[2] * A[2] + A[5] * A[5];
        if (off < 1e-300) break;
#pragma unroll
        for (int pq = 0; pq < 3; ++pq) {
            const int p = (pq == 2) ? 1 : 0;
            const int q = (pq == 0) ? 1 : 2;
            const int r = 3 - p - q;
            double apq = A[p * 3 + q];
            if (apq == 0.0) continue;
            double app = A[p * 3 + p], aqq = A[q * 3 + q];
            double theta = (aqq - app) / (2.0 * apq);
            double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
            double c = 1.0 / sqrt(t * t + 1.0);
            double s = t * c;
            double arp = A[r * 3 + p], arq = A[r * 3 + q];
            A[p * 3 + p] = app - t * apq;
            A[q * 3 + q] = aqq + t * apq;
            A[p * 3 + q] = A[q * 3 + p] = 0.0;
            A[r * 3 + p] = A[p * 3 + r] = c * arp - s * arq;
            A[r * 3 + q] = A[q * 3 + r] = s * arp + c * arq;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                double vkp = V[k * 3 + p], vkq = V[k * 3 + q];
                V[k * 3 + p] = c * vkp - s * vkq;
                V[k * 3 + q] = s * vkp + c * vkq;
            }
        }
    }
    l[0] = A[0];
    l[1] = A[4];
    l[2] = A[8];
}

__device__ __forceinline__ float cof3(const float* m, int i, int j) {
    const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
    return fsub_(fmul_(m[i1 * 3 + j1], m[i2 * 3 + j2]), fmul_(m[i1 * 3 + j2], m[i2 * 3 + j1]));
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Gaussians.h:146-154, 181-201 given the exactly-rounded coordinate sums (s) and centred second moments (acc)
__device__ inline void gaussian_finish(CellStore cs, int g, int n, const double acc[6]) {
    const float den = (float)(n - 1);
    const float cxx = fdiv_((float)acc[0], den), cxy = fdiv_((float)acc[1], den), cxz = fdiv_((float)acc[2], den);
    const float cyy = fdiv_((float)acc[3], den), cyz = fdiv_((float)acc[4], den), czz = fdiv_((float)acc[5], den);
    double A[9] = {cxx, cxy, cxz, cxy, cyy, cyz, cxz, cyz, czz};
    double l[3], V[9];
    jacobi3(A, l, V);
    double lf[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) lf[k] = (double)fmaxf((float)l[k], 0.0001f);  // Gaussians.h:191-194
    float cov[9];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const int rr = r < c ? r : c, cc = r < c ? c : r;
            double v = V[rr * 3 + 0] * lf[0] * V[cc * 3 + 0] + V[rr * 3 + 1] * lf[1] * V[cc * 3 + 1] + V[rr * 3 + 2] * lf[2] * V[cc * 3 + 2];
            cov[r * 3 + c] = (float)v;
        }
    // Gaussians.h:154 cov.inverse(): Eigen fixed-size 3x3 cofactor inverse, float
    float c0 = cof3(cov, 0, 0), c1 = cof3(cov, 1, 0), c2 = cof3(cov, 2, 0);
    float det = fadd_(fmul_(c0, cov[0]), fadd_(fmul_(c1, cov[3]), fmul_(c2, cov[6])));
    float invdet = fdiv_(1.0f, det);
    float* I = cs.info + 9 * (size_t)g;
    I[0] = fmul_(c0, invdet);
    I[1] = fmul_(c1, invdet);
    I[2] = fmul_(c2, invdet);
    I[3] = fmul_(cof3(cov, 0, 1), invdet);
    I[4] = fmul_(cof3(cov, 1, 1), invdet);
    I[7] = fmul_(cof3(cov, 1, 2), invdet);
    I[5] = fmul_(cof3(cov, 2, 1), invdet);
    I[6] = fmul_(cof3(cov, 0, 2), invdet);
    I[8] = fmul_(cof3(cov, 2, 2), invdet);
}

#define GAUSS_WARP_MAX 256
// One warp per accepted set with n <= GAUSS_WARP_MAX members (larger sets: k_gaussian_big over the compact list of
// k_gauss_list): centred second moments -> mom[g][6]; k_gaussian_fin turns them into information matrices.  The member loops are unrolled so that the loads of four strides are in flight together; every
// thread still adds its terms in ascending member order.
__global__ void k_gaussian(const float4* __restrict__ wrec, CellStore cs, const LevelInfo* __restrict__ li, double* __restrict__ mom) {
    DMSA_PDL_ENTER();
    const int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (g >= total_sets(li)) return;
    const int s = cs.start[g], n = cs.n[g];
    if (n > GAUSS_WARP_MAX) return;
    // colwise().mean(): exactly-rounded sum, float division
    double sx = 0, sy = 0, sz = 0;
#pragma unroll 4
    for (int j = lane; j < n; j += 32) {
        float4 p = wrec[s + j];
        sx += (double)p.x;
        sy += (double)p.y;
        sz += (double)p.z;
    }
    sx = warp_sum(sx);
    sy = warp_sum(sy);
    sz = warp_sum(sz);
    const float nf = (float)n;
    const float mx = fdiv_((float)sx, nf), my = fdiv_((float)sy, nf), mz = fdiv_((float)sz, nf);
    double a[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll 4
    for (int j = lane; j < n; j += 32) {
        float4 p = wrec[s + j];
        double cx = (double)fsub_(p.x, mx), cy = (double)fsub_(p.y, my), cz = (double)fsub_(p.z, mz);
        a[0] += cx * cx;
        a[1] += cx * cy;
        a[2] += cx * cz;
        a[3] += cy * cy;
        a[4] += cy * cz;
        a[5] += cz * cz;
    }
#pragma unroll
    for (int k = 0; k < 6; ++k) a[k] = warp_sum(a[k]);
    if (lane == 0)
#pragma unroll
        for (int k = 0; k < 6; ++k) mom[6 * (size_t)g + k] = a[k];
}
// compact list of the sets with n > GAUSS_WARP_MAX (order irrelevant: every set's result depends on the set alone)
__global__ void k_gauss_list(CellStore cs, const LevelInfo* __restrict__ li, int* __restrict__ list, int* __restrict__ count) {
    DMSA_PDL_ENTER();
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= total_sets(li)) return;
    if (cs.n[g] > GAUSS_WARP_MAX) list[atomicAdd(count, 1)] = g;
}
#define GAUSS_BIG_T 1024
// Accepted sets with n > GAUSS_WARP_MAX members: one 1024-thread block per set, a persistent grid strides over the
// compact list.
__global__ void __launch_bounds__(GAUSS_BIG_T) k_gaussian_big(const float4* __restrict__ wrec, CellStore cs, const int* __restrict__ list,
                                                               const int* __restrict__ count, double* __restrict__ mom) {
    DMSA_PDL_ENTER();
    __shared__ double red[GAUSS_BIG_T / 32][6];
    __shared__ float smean[3];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int nbig = *count;
    for (int q = blockIdx.x; q < nbig; q += gridDim.x) {  // block-uniform
        const int g = list[q];
        const int s = cs.start[g], n = cs.n[g];
        __syncthreads();  // shared scratch of the previous set is free
        double sx = 0, sy = 0, sz = 0;
#pragma unroll 4
        for (int j = threadIdx.x; j < n; j += GAUSS_BIG_T) {
            float4 p = wrec[s + j];
            sx += (double)p.x;
            sy += (double)p.y;
            sz += (double)p.z;
        }
        sx = warp_sum(sx);
        sy = warp_sum(sy);
        sz = warp_sum(sz);
        if (lane == 0) {
            red[wid][0] = sx;
            red[wid][1] = sy;
            red[wid][2] = sz;
        }
        __syncthreads();
        if (threadIdx.x < 3) {
            double t = 0;
            for (int w = 0; w < GAUSS_BIG_T / 32; ++w) t += red[w][threadIdx.x];
            smean[threadIdx.x] = fdiv_((float)t, (float)n);
        }
        __syncthreads();
        const float mx = smean[0], my = smean[1], mz = smean[2];
        double a[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll 4
        for (int j = threadIdx.x; j < n; j += GAUSS_BIG_T) {
            float4 p = wrec[s + j];
            double cx = (double)fsub_(p.x, mx), cy = (double)fsub_(p.y, my), cz = (double)fsub_(p.z, mz);
            a[0] += cx * cx;
            a[1] += cx * cy;
            a[2] += cx * cz;
            a[3] += cy * cy;
            a[4] += cy * cz;
            a[5] += cz * cz;
        }
#pragma unroll
        for (int k = 0; k < 6; ++k) a[k] = warp_sum(a[k]);
        __syncthreads();
        if (lane == 0)
            for (int k = 0; k < 6; ++k) red[wid][k] = a[k];
        __syncthreads();
        if (threadIdx.x == 0) {
            double t[6];
            for (int k = 0; k < 6; ++k) {
                t[k] = 0;
                for (int w = 0; w < GAUSS_BIG_T / 32; ++w) t[k] += red[w][k];
            }
            for (int k = 0; k < 6; ++k) mom[6 * (size_t)g + k] = t[k];
        }
    }
}
// Eigen clamp + information matrix + observation weight of every set from its centred second moments: one THREAD per
// set (the 3x3 Jacobi sweeps are long dependent FP64 chains; with one lane per warp they cost 32x the issue slots).
__global__ void k_gaussian_fin(CellStore cs, const LevelInfo* __restrict__ li, const double* __restrict__ mom) {
    DMSA_PDL_ENTER();
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= total_sets(li)) return;
    double a[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) a[k] = mom[6 * (size_t)g + k];
    gaussian_finish(cs, g, cs.n[g], a);
}

// Gaussians.h:172-177: w0 = (1 / n) * observation weight (1, OptimizablePointSet.h:52), w = w0 / mean(w0)   (one block;
// deterministic double reduction, one rounding).  Depends on the set sizes only, so it runs beside the set statistics.
__global__ void k_weights(CellStore cs, const LevelInfo* __restrict__ li) {
    DMSA_PDL_ENTER();
    __shared__ double part[1024];
    const int G = total_sets(li);
    double s = 0;
    for (int g = threadIdx.x; g < G; g += blockDim.x) {
        const float w0 = fmul_(fdiv_(1.0f, (float)cs.n[g]), 1.0f);
        cs.w0[g] = w0;
        s += (double)w0;
    }
    part[threadIdx.x] = s;
    __syncthreads();
    for (int o = blockDim.x / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) part[threadIdx.x] += part[threadIdx.x + o];
        __syncthreads();
    }
    const float mean = fdiv_((float)part[0], (float)G);
    for (int g = threadIdx.x; g < G; g += blockDim.x) cs.w[g] = fdiv_(cs.w0[g], mean);
}

// ---- work decomposition for the cost kernels ------------------------------------------------------------------------
// Sets with at most FUSE_MAX members are evaluated by one block (fused kernel); larger sets are cut into chunks of CH
// members so that no block runs long and the per-set reductions stay parallel.
#define ORDER_CLASSES 32
struct Chunk {
    int cell, start, count, first;  // first: index of the set's first chunk
};
__global__ void k_cell_plan(CellStore cs, const LevelInfo* __restrict__ li, int CH, int fuse_max, int rank, int world, int bound, int* __restrict__ kind,
                            int* __restrict__ nchunk, int* __restrict__ okey, int* __restrict__ oval) {
    DMSA_PDL_ENTER();
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= total_sets(li)) {
        if (g <= bound) nchunk[g] = 0;  // the scan behind this kernel runs over bound + 1 entries
        return;
    }
    const int n = cs.n[g];
    const int k = (g % world == rank) ? (n <= fuse_max ? 1 : 2) : 0;
    kind[g] = k;
    nchunk[g] = (k == 2) ? (n + CH - 1) / CH : 0;
    // size class of the fused kernel's longest-first issue order (class 0 = longest; other ranks' / big sets last)
    const int cls = (k == 1) ? min(ORDER_CLASSES - 2, (fuse_max - n) * (ORDER_CLASSES - 1) / (fuse_max + 1)) : ORDER_CLASSES - 1;
    okey[g] = cls;
    atomicAdd(&oval[cls], 1);  // class histogram (oval[0..ORDER_CLASSES) must be zeroed before the launch)
}
// order[] = set indices grouped by size class, longest class first.  The position inside a class comes from an atomic
// cursor: the order is a scheduling hint only and does not influence any result.
__global__ void k_cell_order(const LevelInfo* __restrict__ li, const int* __restrict__ okey, int* __restrict__ hist_cursor, int* __restrict__ order) {
    DMSA_PDL_ENTER();
    __shared__ int base[ORDER_CLASSES];
    if (threadIdx.x == 0) {
        int acc = 0;
        for (int c = 0; c < ORDER_CLASSES; ++c) {
            base[c] = acc;
            acc += hist_cursor[c];
        }
    }
    __syncthreads();
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= total_sets(li)) return;
    const int c = okey[g];
    const int pos = atomicAdd(&hist_cursor[ORDER_CLASSES + c], 1);
    order[base[c] + pos] = g;
}
// chunk_off = exclusive scan of nchunk (G+1 entries, last = total)
__global__ void k_chunk_fill(CellStore cs, const LevelInfo* __restrict__ li, int CH, const int* __restrict__ nchunk, const int* __restrict__ chunk_off,
                             Chunk* __restrict__ chunks) {
    DMSA_PDL_ENTER();
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= total_sets(li)) return;
    const int nc = nchunk[g], o = chunk_off[g], s = cs.start[g], n = cs.n[g];
    for (int c = 0; c < nc; ++c) {
        Chunk ch;
        ch.cell = g;
        ch.start = s + c * CH;
        ch.count = min(CH, n - c * CH);
        ch.first = o;
        chunks[o + c] = ch;
    }
}

}  // namespace dmsa
