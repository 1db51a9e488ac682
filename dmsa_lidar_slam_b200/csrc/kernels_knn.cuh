// kernels_knn.cuh — SURVEY §8(f) rank 2: the nearest-neighbour work right before the sliding-window pass.
//   DmsaSlam.h:264-345  addStaticPoints: a keyframe point becomes a static map point iff its nearest window point is within
//                       minGridSize (PCL KdTreeFLANN, flann::L2_Simple<float> squared distance) and it is visible (:347-363)
//   DmsaSlam.h:377-414  getOverlap: fraction of the window points with an active map point within maxDist
// "nearest squared distance <= threshold" == "some point has squared distance <= threshold": an exact, order-free decision.
// A uniform grid (cell edge = radius * 1.000001, cell of a point = floor(double(x) / h)) is built over the searched cloud by
// sorting 63-bit cell keys; a query walks the 27 neighbouring cells (binary search of the key in the sorted array).  The
// float arithmetic is FLANN's: ((dx*dx + dy*dy) + dz*dz) with separate roundings (no FMA).
#pragma once
#include <cuda_runtime.h>

#include "kernels_sets.cuh"

namespace dmsa {

__device__ __forceinline__ unsigned long long grid_key(long long cx, long long cy, long long cz) {
    return (unsigned long long)(cx + (1 << 20)) | ((unsigned long long)(cy + (1 << 20)) << 21) | ((unsigned long long)(cz + (1 << 20)) << 42);
}
__device__ __forceinline__ bool finite3(float x, float y, float z) { return isfinite(x) && isfinite(y) && isfinite(z); }

// keys of the searched cloud; non-finite points get the all-ones key (sorted last, never matched)
__global__ void k_grid_keys(const float4* __restrict__ pts, int n, double h, unsigned long long* __restrict__ keys, int* __restrict__ idx) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 p = pts[i];
    unsigned long long k = ~0ull;
    if (finite3(p.x, p.y, p.z)) k = grid_key((long long)floor((double)p.x / h), (long long)floor((double)p.y / h), (long long)floor((double)p.z / h));
    keys[i] = k;
    idx[i] = i;
}
__global__ void k_grid_gather(const float4* __restrict__ pts, const int* __restrict__ sidx, int n, float4* __restrict__ spts) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) spts[i] = pts[sidx[i]];
}

struct GridView {
    const unsigned long long* keys;  // sorted
    const float4* pts;               // in key order
    int n;
    double h;
};
__device__ __forceinline__ bool grid_any_within(const GridView& g, float qx, float qy, float qz, float max_sq) {
    if (!finite3(qx, qy, qz)) return false;
    const long long cx = (long long)floor((double)qx / g.h), cy = (long long)floor((double)qy / g.h), cz = (long long)floor((double)qz / g.h);
    for (int dz = -1; dz <= 1; ++dz)
        for (int dy = -1; dy <= 1; ++dy)
            for (int dx = -1; dx <= 1; ++dx) {
                const unsigned long long k = grid_key(cx + dx, cy + dy, cz + dz);
                int lo = 0, hi = g.n;
                while (lo < hi) {
                    const int mid = (lo + hi) >> 1;
                    if (__ldg(g.keys + mid) < k)
                        lo = mid + 1;
                    else
                        hi = mid;
                }
                for (int i = lo; i < g.n && __ldg(g.keys + i) == k; ++i) {
                    const float4 p = __ldg(g.pts + i);
                    const float ex = fsub_(qx, p.x), ey = fsub_(qy, p.y), ez = fsub_(qz, p.z);
                    const float d2 = fadd_(fadd_(fmul_(ex, ex), fmul_(ey, ey)), fmul_(ez, ez));  // flann::L2_Simple<float>
                    if (d2 <= max_sq) return true;
                }
            }
    return false;
}

// addStaticPoints inner loop (DmsaSlam.h:304-339) for one keyframe cloud of pcl::PointNormal (3 float4 per point)
__global__ void k_select_static(GridView g, const float4* __restrict__ cloud, int n, float px, float py, float pz, float max_sq,
                                unsigned char* __restrict__ selected, int* __restrict__ count) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    bool sel = false;
    if (j < n) {
        const float4 p = cloud[3 * (size_t)j], nr = cloud[3 * (size_t)j + 1];
        if (g.n > 0 && grid_any_within(g, p.x, p.y, p.z, max_sq)) {
            // isVisible :347-363: d = p . n, res = pos . n - d (Eigen 3-vector dot: a0 + (a1 + a2)), res >= -0.00001
            const float d = fadd_(fmul_(p.x, nr.x), fadd_(fmul_(p.y, nr.y), fmul_(p.z, nr.z)));
            const float e = fadd_(fmul_(px, nr.x), fadd_(fmul_(py, nr.y), fmul_(pz, nr.z)));
            sel = (double)fsub_(e, d) >= -0.00001;
        }
        selected[j] = sel ? 1 : 0;
    }
    const unsigned m = __ballot_sync(0xffffffffu, sel);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(count, __popc(m));
}
// getOverlap (DmsaSlam.h:377-414): window points with a point of the searched cloud within max_dist
__global__ void k_overlap_count(GridView g, const float4* __restrict__ window, int n, float max_sq, int* __restrict__ count) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    bool hit = false;
    if (j < n) {
        const float4 q = window[j];
        hit = grid_any_within(g, q.x, q.y, q.z, max_sq);
    }
    const unsigned m = __ballot_sync(0xffffffffu, hit);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(count, __popc(m));
}

}  // namespace dmsa
