// kernels_pre.cuh — SURVEY §8(f) rank 3: the scan pre-processing in front of the optimizer.
//   helpers.h:67-182    randomGridDownsampling: one member of every PCL octree leaf, leaves in depth-first order, member
//                       id = int(r * (n - 1)) with r = rand() / RAND_MAX drawn once per leaf in that order
//   DmsaSlam.h:570-634  preProcess: adaptive grid size (0.4 / 0.3 / 0.2 / 0.15), range cut at the max_num-th smallest range,
//                       pcl::transformPointCloud with lidarToImuTform, w = 1
// The octree (lattice anchor, keys, root growth replay, Morton leaf order) is the one of the set build (kernels_sets.cuh,
// kernels_sort.cuh); the random numbers are glibc's rand() sequence for a caller-supplied seed, generated on the host
// (dmsa_b200_pre.inl) while the device builds the octree.
#pragma once
#include <cuda_runtime.h>

#include "pdl.cuh"

#include "kernels_sets.cuh"

namespace dmsa {

// xyz of a point record (stride bytes apart, three floats at offset 0) -> float4 (w = 1)
__global__ void k_pre_unpack(const unsigned char* __restrict__ raw, int n, int stride, float4* __restrict__ pts) {
    DMSA_PDL_ENTER();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* p = reinterpret_cast<const float*>(raw + (size_t)i * stride);
    pts[i] = make_float4(p[0], p[1], p[2], 1.0f);
}
// helpers.h:96-103: the member of leaf c (members in ascending point index == PCL's container order)
__global__ void k_pre_pick(const int* __restrict__ raw_start, const int* __restrict__ sidx, const int* __restrict__ rnd, int R, int* __restrict__ pick) {
    DMSA_PDL_ENTER();
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= R) return;
    const int s = raw_start[c], n = raw_start[c + 1] - s;
    const double r = (double)rnd[c] / 2147483647.0;    // ((double)rand() / (RAND_MAX))
    const int id = (int)(r * (double)(n - 1));         // static_cast<int>(r * (double)(indices.size() - 1))
    pick[c] = sidx[s + id];
}
// ranges of the picked points: Eigen::Vector3f(x, y, z).norm() = sqrt(x^2 + (y^2 + z^2))   (DmsaSlam.h:599-603); key = float bits
__global__ void k_pre_ranges(const float4* __restrict__ pts, const int* __restrict__ pick, int R, float* __restrict__ range,
                             unsigned long long* __restrict__ key) {
    DMSA_PDL_ENTER();
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= R) return;
    const float4 p = pts[pick[c]];
    const float r = __fsqrt_rn(fadd_(fmul_(p.x, p.x), fadd_(fmul_(p.y, p.y), fmul_(p.z, p.z))));
    range[c] = r;
    key[c] = (unsigned long long)__float_as_uint(r);  // ranges are >= 0 and finite: the bit pattern orders like the value
}
// thresRange = max(rangesSorted[min(max_num, size - 1)], minDistDS) (:609); keep[c] = ranges[c] < thresRange && ranges[c] > min_dist (:616)
__global__ void k_pre_keep(const float* __restrict__ range, const unsigned long long* __restrict__ sorted_key, int R, int max_num, float min_dist_ds,
                           float min_dist, int* __restrict__ keep) {
    DMSA_PDL_ENTER();
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= R) return;
    const float kth = __uint_as_float((unsigned)sorted_key[min(max_num, R - 1)]);
    const float thres = fmaxf(kth, min_dist_ds);
    const float r = range[c];
    keep[c] = (r < thres && r > min_dist) ? 1 : 0;
}
// pos = exclusive scan of keep.  Output record = the raw record with xyz transformed like pcl::transformPointCloud (PCL 1.10
// common/impl/transforms.hpp, pcl::detail::Transformer<float>::se3, SSE2 build: p0 + (p1 + (p2 + c3)) per component with
// p_i = src[i] * column_i), then data[3] = 1 (:629-630).  T: column-major Matrix4f.
struct PreTform {
    float m[16];
};
__global__ void k_pre_emit(const unsigned char* __restrict__ raw, int stride, const int* __restrict__ pick, const int* __restrict__ keep,
                           const int* __restrict__ pos, int R, PreTform T, unsigned char* __restrict__ out, int* __restrict__ n_out) {
    DMSA_PDL_ENTER();
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= R) return;
    if (c == R - 1) *n_out = pos[c] + keep[c];
    if (!keep[c]) return;
    const uint4* src = reinterpret_cast<const uint4*>(raw + (size_t)pick[c] * stride);
    uint4* dst = reinterpret_cast<uint4*>(out + (size_t)pos[c] * stride);
    const uint4 a = src[0];
    const float x = __uint_as_float(a.x), y = __uint_as_float(a.y), z = __uint_as_float(a.z);
    float o[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) o[r] = fadd_(fmul_(x, T.m[r]), fadd_(fmul_(y, T.m[4 + r]), fadd_(fmul_(z, T.m[8 + r]), T.m[12 + r])));
    dst[0] = make_uint4(__float_as_uint(o[0]), __float_as_uint(o[1]), __float_as_uint(o[2]), __float_as_uint(1.0f));
    for (int q = 1; q < stride / 16; ++q) dst[q] = src[q];
}

}  // namespace dmsa
