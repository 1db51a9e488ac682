// kernels_knn.cuh — nearest-neighbour work around the optimizer.
//   DmsaSlam.h:264-345  addStaticPoints: a keyframe point becomes a static map point iff its nearest window point is within
//                       minGridSize (PCL KdTreeFLANN, flann::L2_Simple<float> squared distance) and it is visible (:347-363)
//   DmsaSlam.h:377-414  getOverlap: fraction of the window points with an active map point within maxDist
//   DmsaSlam.h:557-568  updateNormals: pcl::NormalEstimationOMP with setKSearch(6) (PCL 1.10, not part of /root/reference:
//                       features/normal_3d.h computePointNormal, common/centroid.hpp computeMeanAndCovarianceMatrix,
//                       common/eigen.hpp eigen33 / computeRoots, flipNormalTowardsViewpoint — restated from the published code)
//
// One structure serves all of them: a HASHED uniform grid.  cell = floor(double(x) / h) per axis, bucket = hash(cell) & (B - 1);
// the points are counting-sorted by bucket (histogram with atomics, one chained scan, scatter with atomic cursors).  A query
// visits the buckets of the cells it needs and tests EVERY point it finds there with the exact float distance
// ((dx*dx + dy*dy) + dz*dz, separate roundings == flann::L2_Simple<float>), so hash collisions only add candidates and the
// arbitrary order inside a bucket cannot change a result: radius decisions are order-free, and the k nearest neighbours are
// selected by the total order (distance, point index).
#pragma once
#include <cuda_runtime.h>

#include "pdl.cuh"

#include "kernels_sets.cuh"

namespace dmsa {

__device__ __forceinline__ bool finite3(float x, float y, float z) { return isfinite(x) && isfinite(y) && isfinite(z); }

struct HashGrid {
    const int* start;   // [B + 1] first slot of every bucket
    const float4* pts;  // points in bucket order: xyz, w = original index (int bits)
    int n, B;           // B: power of two
    double h;           // cell edge
};
__device__ __forceinline__ int grid_cell(float x, double h) { return (int)floor((double)x / h); }
__device__ __forceinline__ unsigned grid_bucket(int cx, int cy, int cz, int B) {
    unsigned hsh = (unsigned)cx * 73856093u ^ (unsigned)cy * 19349663u ^ (unsigned)cz * 83492791u;
    hsh ^= hsh >> 15;
    return hsh & (unsigned)(B - 1);
}
// bucket of every point (-1: not finite, in no bucket) + bucket histogram
__global__ void k_hg_count(const float4* __restrict__ pts, int n, int stride4, double h, int B, int* __restrict__ bucket, int* __restrict__ counts) {
    DMSA_PDL_ENTER();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 p = pts[(size_t)i * stride4];
    int b = -1;
    if (finite3(p.x, p.y, p.z)) {
        b = (int)grid_bucket(grid_cell(p.x, h), grid_cell(p.y, h), grid_cell(p.z, h), B);
        atomicAdd(counts + b, 1);
    }
    bucket[i] = b;
}
__global__ void k_hg_fill(const float4* __restrict__ pts, int n, int stride4, const int* __restrict__ bucket, const int* __restrict__ start,
                          int* __restrict__ cursor, float4* __restrict__ spts) {
    DMSA_PDL_ENTER();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int b = bucket[i];
    if (b < 0) return;
    float4 p = pts[(size_t)i * stride4];
    p.w = __int_as_float(i);
    spts[start[b] + atomicAdd(cursor + b, 1)] = p;
}

__device__ __forceinline__ float l2_simple(float qx, float qy, float qz, const float4& p) {  // flann::L2_Simple<float>, dimension 3
    const float ex = fsub_(qx, p.x), ey = fsub_(qy, p.y), ez = fsub_(qz, p.z);
    return fadd_(fadd_(fmul_(ex, ex), fmul_(ey, ey)), fmul_(ez, ez));
}
// cell edge = radius * 1.000001: every point within `radius` of the query lies in the 27 cells around the query's cell
__device__ __forceinline__ bool grid_any_within(const HashGrid& g, float qx, float qy, float qz, float max_sq) {
    if (!finite3(qx, qy, qz)) return false;
    const int cx = grid_cell(qx, g.h), cy = grid_cell(qy, g.h), cz = grid_cell(qz, g.h);
    for (int dz = -1; dz <= 1; ++dz)
        for (int dy = -1; dy <= 1; ++dy)
            for (int dx = -1; dx <= 1; ++dx) {
                const unsigned b = grid_bucket(cx + dx, cy + dy, cz + dz, g.B);
                const int s = __ldg(g.start + b), e = __ldg(g.start + b + 1);
                for (int i = s; i < e; ++i)
                    if (l2_simple(qx, qy, qz, __ldg(g.pts + i)) <= max_sq) return true;
            }
    return false;
}

// addStaticPoints inner loop (DmsaSlam.h:304-339) for one keyframe cloud of pcl::PointNormal (3 float4 per point)
__global__ void k_select_static(HashGrid g, const float4* __restrict__ cloud, int n, float px, float py, float pz, float max_sq,
                                unsigned char* __restrict__ selected, int* __restrict__ count) {
    DMSA_PDL_ENTER();
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
__global__ void k_overlap_count(HashGrid g, const float4* __restrict__ window, int n, float max_sq, int* __restrict__ count) {
    DMSA_PDL_ENTER();
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    bool hit = false;
    if (j < n) {
        const float4 q = window[j];
        hit = grid_any_within(g, q.x, q.y, q.z, max_sq);
    }
    const unsigned m = __ballot_sync(0xffffffffu, hit);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(count, __popc(m));
}

// ---- k nearest neighbours (k = KNN_K = 6, the query point itself included) ----------------------------------------------
// Candidates are gathered shell by shell (Chebyshev radius rho around the query's cell).  After shell rho every point not
// yet seen lies outside the cube of (2 rho + 1)^3 cells, i.e. at least `edge` away from the query, where edge = distance
// from the query to the nearest face of that cube (exact in double).  The search stops once the k-th candidate is closer
// than edge (with a relative margin of 1e-6 for the rounding of the float distances); queries that are still open after
// KNN_MAX_RHO shells scan the whole cloud.  Result: indices in ascending (distance, index) order — FLANN's result order
// for pairwise different distances.
#define KNN_K 6
#define KNN_MAX_RHO 6
struct Knn6 {
    float d[KNN_K];
    int i[KNN_K];
};
__device__ __forceinline__ void knn_init(Knn6& r) {
#pragma unroll
    for (int k = 0; k < KNN_K; ++k) {
        r.d[k] = 3.402823466e+38f;
        r.i[k] = 0x7fffffff;
    }
}
__device__ __forceinline__ void knn_push(Knn6& r, float d, int i) {
    if (!(d < r.d[KNN_K - 1] || (d == r.d[KNN_K - 1] && i < r.i[KNN_K - 1]))) return;
    // a point can be met twice: once in its own cell's bucket, once in the bucket of another visited cell that collides with it
#pragma unroll
    for (int k = 0; k < KNN_K; ++k)
        if (r.i[k] == i) return;
    r.d[KNN_K - 1] = d;
    r.i[KNN_K - 1] = i;
#pragma unroll
    for (int k = KNN_K - 1; k > 0; --k) {
        const bool sw = r.d[k] < r.d[k - 1] || (r.d[k] == r.d[k - 1] && r.i[k] < r.i[k - 1]);
        if (sw) {
            const float td = r.d[k];
            r.d[k] = r.d[k - 1];
            r.d[k - 1] = td;
            const int ti = r.i[k];
            r.i[k] = r.i[k - 1];
            r.i[k - 1] = ti;
        }
    }
}
__device__ __forceinline__ void knn_bucket(const HashGrid& g, int cx, int cy, int cz, float qx, float qy, float qz, Knn6& r) {
    const unsigned b = grid_bucket(cx, cy, cz, g.B);
    const int s = __ldg(g.start + b), e = __ldg(g.start + b + 1);
    for (int i = s; i < e; ++i) {  // (points of colliding cells included: exact distances, duplicates dropped by knn_push)
        const float4 p = __ldg(g.pts + i);
        knn_push(r, l2_simple(qx, qy, qz, p), __float_as_int(p.w));
    }
}
__device__ inline void knn6_query(const HashGrid& g, float qx, float qy, float qz, Knn6& r) {
    knn_init(r);
    const int cx = grid_cell(qx, g.h), cy = grid_cell(qy, g.h), cz = grid_cell(qz, g.h);
    // distance from the query to the nearer face of its own cell, per axis (exact: the cell bounds are c h and (c + 1) h)
    double inner = g.h;
    {
        const double q[3] = {(double)qx, (double)qy, (double)qz};
        const int c[3] = {cx, cy, cz};
#pragma unroll
        for (int a = 0; a < 3; ++a) inner = fmin(inner, fmin(q[a] - (double)c[a] * g.h, ((double)c[a] + 1.0) * g.h - q[a]));
        inner = fmax(inner, 0.0);
    }
    for (int rho = 0; rho <= KNN_MAX_RHO; ++rho) {
        for (int dz = -rho; dz <= rho; ++dz)
            for (int dy = -rho; dy <= rho; ++dy) {
                const bool face = (dz == -rho || dz == rho || dy == -rho || dy == rho);
                if (face) {
                    for (int dx = -rho; dx <= rho; ++dx) knn_bucket(g, cx + dx, cy + dy, cz + dz, qx, qy, qz, r);
                } else {
                    knn_bucket(g, cx - rho, cy + dy, cz + dz, qx, qy, qz, r);
                    knn_bucket(g, cx + rho, cy + dy, cz + dz, qx, qy, qz, r);
                }
            }
        const double edge = ((double)rho * g.h + inner) * (1.0 - 1e-6);
        if (r.i[KNN_K - 1] != 0x7fffffff && (double)r.d[KNN_K - 1] < edge * edge) return;
    }
    // open after the last shell (isolated point): exhaustive scan
    knn_init(r);
    const int filled = __ldg(g.start + g.B);  // finite points only
    for (int i = 0; i < filled; ++i) {
        const float4 p = __ldg(g.pts + i);
        knn_push(r, l2_simple(qx, qy, qz, p), __float_as_int(p.w));
    }
}

// ---- PCL 1.10 surface normal of a k-neighbourhood ---------------------------------------------------------------------
// computeMeanAndCovarianceMatrix (common/impl/centroid.hpp, dense cloud): ONE pass of float accumulators over the
// neighbours in search-result order, accu /= n, covariance = E[xy] - E[x]E[y]
// eigen33 (common/impl/eigen.hpp): scale by the largest |entry|, closed-form roots (computeRoots), eigenvector of the
// smallest root = the largest of the three row cross products of (A - lambda I)
__device__ __forceinline__ void pcl_roots2(float b, float c, float roots[3]) {  // computeRoots2
    roots[0] = 0.0f;
    float d = (float)((double)fmul_(b, b) - 4.0 * (double)c);  // Scalar (b * b - 4.0 * c): the subtraction runs in double
    if (d < 0.0f) d = 0.0f;
    const float sd = __fsqrt_rn(d);
    roots[2] = fmul_(0.5f, fadd_(b, sd));
    roots[1] = fmul_(0.5f, fsub_(b, sd));
}
__device__ inline void pcl_roots(const float m[9], float roots[3]) {  // computeRoots
    // c0 = m00 m11 m22 + 2 m01 m02 m12 - m00 m12^2 - m11 m02^2 - m22 m01^2   (left to right)
    const float m00 = m[0], m01 = m[1], m02 = m[2], m11 = m[4], m12 = m[5], m22 = m[8];
    float c0 = fmul_(fmul_(m00, m11), m22);
    c0 = fadd_(c0, fmul_(fmul_(fmul_(2.0f, m01), m02), m12));
    c0 = fsub_(c0, fmul_(fmul_(m00, m12), m12));
    c0 = fsub_(c0, fmul_(fmul_(m11, m02), m02));
    c0 = fsub_(c0, fmul_(fmul_(m22, m01), m01));
    float c1 = fsub_(fmul_(m00, m11), fmul_(m01, m01));
    c1 = fadd_(c1, fmul_(m00, m22));
    c1 = fsub_(c1, fmul_(m02, m02));
    c1 = fadd_(c1, fmul_(m11, m22));
    c1 = fsub_(c1, fmul_(m12, m12));
    const float c2 = fadd_(fadd_(m00, m11), m22);
    if (fabsf(c0) < 1.1920928955078125e-07f) {  // one root is 0 -> quadratic equation
        pcl_roots2(c2, c1, roots);
        return;
    }
    const float s_inv3 = (float)(1.0 / 3.0);
    const float s_sqrt3 = __fsqrt_rn(3.0f);
    const float c2_over_3 = fmul_(c2, s_inv3);
    float a_over_3 = fmul_(fsub_(c1, fmul_(c2, c2_over_3)), s_inv3);
    if (a_over_3 > 0.0f) a_over_3 = 0.0f;
    const float half_b = fmul_(0.5f, fadd_(c0, fmul_(c2_over_3, fsub_(fmul_(fmul_(2.0f, c2_over_3), c2_over_3), c1))));
    float q = fadd_(fmul_(half_b, half_b), fmul_(fmul_(a_over_3, a_over_3), a_over_3));
    if (q > 0.0f) q = 0.0f;
    const float rho = __fsqrt_rn(-a_over_3);
    const float theta = fmul_(atan2f(__fsqrt_rn(-q), half_b), s_inv3);
    const float cos_theta = cosf(theta), sin_theta = sinf(theta);
    roots[0] = fadd_(c2_over_3, fmul_(fmul_(2.0f, rho), cos_theta));
    roots[1] = fsub_(c2_over_3, fmul_(rho, fadd_(cos_theta, fmul_(s_sqrt3, sin_theta))));
    roots[2] = fsub_(c2_over_3, fmul_(rho, fsub_(cos_theta, fmul_(s_sqrt3, sin_theta))));
    // sort in increasing order
    if (roots[0] >= roots[1]) {
        const float t = roots[0];
        roots[0] = roots[1];
        roots[1] = t;
    }
    if (roots[1] >= roots[2]) {
        float t = roots[1];
        roots[1] = roots[2];
        roots[2] = t;
        if (roots[0] >= roots[1]) {
            t = roots[0];
            roots[0] = roots[1];
            roots[1] = t;
        }
    }
    if (roots[0] <= 0.0f) pcl_roots2(c2, c1, roots);  // an eigenvalue of a positive semi-definite matrix cannot be negative
}
__device__ __forceinline__ void cross3f(const float a[3], const float b[3], float o[3]) {  // Eigen cross(): a1 b2 - a2 b1, ...
    o[0] = fsub_(fmul_(a[1], b[2]), fmul_(a[2], b[1]));
    o[1] = fsub_(fmul_(a[2], b[0]), fmul_(a[0], b[2]));
    o[2] = fsub_(fmul_(a[0], b[1]), fmul_(a[1], b[0]));
}
__device__ __forceinline__ float sqnorm3f(const float a[3]) { return fadd_(fmul_(a[0], a[0]), fadd_(fmul_(a[1], a[1]), fmul_(a[2], a[2]))); }
// normal (nx, ny, nz) and curvature of the neighbourhood nb[0..cnt); false: fewer than 3 neighbours (PCL writes NaN)
__device__ inline bool pcl_point_normal(const float4* nb, int cnt, float n[3], float& curvature) {
    if (cnt < 3) return false;
    float accu[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int k = 0; k < cnt; ++k) {
        const float x = nb[k].x, y = nb[k].y, z = nb[k].z;
        accu[0] = fadd_(accu[0], fmul_(x, x));
        accu[1] = fadd_(accu[1], fmul_(x, y));
        accu[2] = fadd_(accu[2], fmul_(x, z));
        accu[3] = fadd_(accu[3], fmul_(y, y));
        accu[4] = fadd_(accu[4], fmul_(y, z));
        accu[5] = fadd_(accu[5], fmul_(z, z));
        accu[6] = fadd_(accu[6], x);
        accu[7] = fadd_(accu[7], y);
        accu[8] = fadd_(accu[8], z);
    }
    const float nf = (float)cnt;
#pragma unroll
    for (int k = 0; k < 9; ++k) accu[k] = fdiv_(accu[k], nf);
    float cov[9];
    cov[0] = fsub_(accu[0], fmul_(accu[6], accu[6]));
    cov[1] = fsub_(accu[1], fmul_(accu[6], accu[7]));
    cov[2] = fsub_(accu[2], fmul_(accu[6], accu[8]));
    cov[4] = fsub_(accu[3], fmul_(accu[7], accu[7]));
    cov[5] = fsub_(accu[4], fmul_(accu[7], accu[8]));
    cov[8] = fsub_(accu[5], fmul_(accu[8], accu[8]));
    cov[3] = cov[1];
    cov[6] = cov[2];
    cov[7] = cov[5];
    // eigen33 (smallest eigenvalue and its eigenvector)
    float scale = 0.0f;
#pragma unroll
    for (int k = 0; k < 9; ++k) scale = fmaxf(scale, fabsf(cov[k]));
    if (scale <= 1.17549435e-38f) scale = 1.0f;
    float sm[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) sm[k] = fdiv_(cov[k], scale);
    float roots[3];
    pcl_roots(sm, roots);
    const float eigenvalue = fmul_(roots[0], scale);
    sm[0] = fsub_(sm[0], roots[0]);
    sm[4] = fsub_(sm[4], roots[0]);
    sm[8] = fsub_(sm[8], roots[0]);
    float v1[3], v2[3], v3[3];
    cross3f(sm, sm + 3, v1);
    cross3f(sm, sm + 6, v2);
    cross3f(sm + 3, sm + 6, v3);
    const float l1 = sqnorm3f(v1), l2 = sqnorm3f(v2), l3 = sqnorm3f(v3);
    const float* v;
    float len;
    if (l1 >= l2 && l1 >= l3) {
        v = v1;
        len = l1;
    } else if (l2 >= l1 && l2 >= l3) {
        v = v2;
        len = l2;
    } else {
        v = v3;
        len = l3;
    }
    const float sl = __fsqrt_rn(len);
    n[0] = fdiv_(v[0], sl);
    n[1] = fdiv_(v[1], sl);
    n[2] = fdiv_(v[2], sl);
    // solvePlaneParameters: curvature = |lambda_0 / trace(cov)|
    const float eig_sum = fadd_(fadd_(cov[0], cov[4]), cov[8]);
    curvature = eig_sum != 0.0f ? fabsf(fdiv_(eigenvalue, eig_sum)) : 0.0f;
    return true;
}

// NormalEstimationOMP::computeFeature for every point of a pcl::PointNormal cloud (3 float4 per point), k = 6, search
// surface = the cloud itself; normals flipped towards the view point (flipNormalTowardsViewpoint, float& overload).
// nn_out (optional): the 6 neighbour indices of every point in result order.
__global__ void __launch_bounds__(128) k_normals_knn6(HashGrid g, float4* __restrict__ cloud, int n, float vx, float vy, float vz, int* __restrict__ nn_out) {
    DMSA_PDL_ENTER();
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const float4 q = cloud[3 * (size_t)j];
    float4 nr = make_float4(__int_as_float(0x7fc00000), __int_as_float(0x7fc00000), __int_as_float(0x7fc00000), 0.0f);
    float curv = __int_as_float(0x7fc00000);
    Knn6 r;
    knn_init(r);
    if (finite3(q.x, q.y, q.z)) {  // (PCL: a non-finite query point gets NaN normals)
        knn6_query(g, q.x, q.y, q.z, r);
        float4 nb[KNN_K];
        int cnt = 0;
#pragma unroll
        for (int k = 0; k < KNN_K; ++k)
            if (r.i[k] != 0x7fffffff) nb[cnt++] = cloud[3 * (size_t)r.i[k]];
        float nv[3], c;
        if (pcl_point_normal(nb, cnt, nv, c)) {
            const float wx = fsub_(vx, q.x), wy = fsub_(vy, q.y), wz = fsub_(vz, q.z);
            const float cos_theta = fadd_(fadd_(fmul_(wx, nv[0]), fmul_(wy, nv[1])), fmul_(wz, nv[2]));
            if (cos_theta < 0.0f) {
                nv[0] = fmul_(nv[0], -1.0f);
                nv[1] = fmul_(nv[1], -1.0f);
                nv[2] = fmul_(nv[2], -1.0f);
            }
            nr = make_float4(nv[0], nv[1], nv[2], 0.0f);
            curv = c;
        }
    }
    cloud[3 * (size_t)j + 1] = nr;
    float4 t = cloud[3 * (size_t)j + 2];
    t.x = curv;
    cloud[3 * (size_t)j + 2] = t;
    if (nn_out)
#pragma unroll
        for (int k = 0; k < KNN_K; ++k) nn_out[(size_t)KNN_K * j + k] = r.i[k] == 0x7fffffff ? -1 : r.i[k];
}

}  // namespace dmsa
