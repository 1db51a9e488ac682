// kernels_pose.cuh — pose parameters -> per-parameter-vector transform tables (double math, float tables).
//
// For a batch of V parameter vectors (base + P forward-difference vectors, or the 9 line-search
// vectors) these kernels restate, once per vector:
//   Poses.h:72-76                     setParamsFromVector (parameter layout [w_1..w_{n-1} | t_1..t_{n-1}])
//   ConsecutivePoses.h:26-43          relative2global
//   ContinuousTrajectory.h:189-226    updateTrajDenseTforms (slerp + barycentric-rational + axang2rotm -> Matrix4f)
//   MapManagement.h:133-138           per-keyframe Matrix4f
//   ContinuousTrajectory.h:603-663    updateImuError          (extra residual rows)
//   MapManagement.h:210-252           gravity / odometry rows (extra residual rows)
// Table layout in HBM: Mtab[(row * Vld + v) * 12 + 4*r + c], rows 0..2 of the Matrix4f, so that the 32
// lanes of a warp (32 consecutive v) read 32 consecutive 48-byte records.
#pragma once
#include <cuda_runtime.h>

#include "pdl.cuh"

#include "se3_math.cuh"

namespace dmsa {

struct PoseBatch {
    int V;                 // parameter vectors in the batch
    int Vld;               // padded V (multiple of 32)
    int P;                 // parameters per vector
    int n;                 // control poses / keyframes
    const double* params;  // [V][P]
    double pose0[6];       // relative pose 0 (orientation, translation): not a parameter (Poses.h:64-76)
    double* globO_t;       // [(k*3+a)*Vld + v]
    double* globT_t;       // [(k*3+a)*Vld + v]
    double* quat_t;        // [(k*4+c)*Vld + v]
    double* extra;         // [E][Vld] additional residual rows (may be null)
    float* Mpair;          // pair-interleaved second copy of the table (may be null)
};

struct TrajTiming {
    int n_total;
    const int* seg;        // [n_total] rightIndex of getInterpRotation (ContinuousTrajectory.h:573-575)
    const double* urel;    // [n_total] t_rel (:578-581)
    const double* fh;      // [n_total][n] w_i / (t_j - s_i)   (Boost barycentric_rational::operator())
    const int* hit;        // [n_total] index i with t_j == s_i, else -1
};

struct ImuFactors {       // ContinuousTrajectory.h:36-43, 520-553
    int enabled;
    const double* preRot;  // [n][9]
    const double* prePos;  // [n][3]
    const double* preVel;  // [n][3]
    const double* covInv;  // [n][81]
    const int* paramIdx;   // [n]
    const double* stamps;  // [n]
    double balancing, dt_res, gravity[3];
};

struct KfFactors {        // MapManagement.h:36-70
    int useGrav, useOdom;
    const double* measGrav;  // [n][3]
    const int* plausible;    // [n]
    const double* odomT;     // [n][3]
    const double* odomR;     // [n][9]
    double balanceGrav, balanceOdom, gravity[3];
};

// Builds the forward-difference batch: row 0 = p, row k+1 = p + h e_k   (DmsaOptimizer.h:209-218)
__global__ void k_make_fd_batch(const double* __restrict__ p, int P, double h, double* __restrict__ out) {
    DMSA_PDL_ENTER();
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int total = (P + 1) * P;
    if (i >= total) return;
    int v = i / P, k = i % P;
    double x = p[k];
    if (v == k + 1) x += h;
    out[i] = x;
}
// Builds the line-search batch: row k-1 = p + 0.1*k*step, k = 1..9   (DmsaOptimizer.h:160-162)
__global__ void k_make_ls_batch(const double* __restrict__ p, const double* __restrict__ step, int P, double* __restrict__ out) {
    DMSA_PDL_ENTER();
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 9 * P) return;
    int v = i / P, k = i % P;
    out[i] = p[k] + 0.1 * (double)(v + 1) * step[k];
}

// second copy of a table entry in the pair-interleaved layout read by the pair-packed cost kernels:
// Mpair[((row * Vld/2 + v/2) * 12 + c) * 2 + (v & 1)] = M[row][v][c]
__device__ __forceinline__ void store_pair_entry(float* __restrict__ Mpair, int row, int v, int Vld, const float4& r0, const float4& r1, const float4& r2) {
    if (Mpair == nullptr) return;
    float* b = Mpair + ((size_t)row * (Vld >> 1) + (v >> 1)) * 24 + (v & 1);
    b[0] = r0.x; b[2] = r0.y; b[4] = r0.z; b[6] = r0.w;
    b[8] = r1.x; b[10] = r1.y; b[12] = r1.z; b[14] = r1.w;
    b[16] = r2.x; b[18] = r2.y; b[20] = r2.z; b[22] = r2.w;
}

// One block per parameter vector.  Shared memory: n * (9 + 9 + 3 + 3) doubles.
template <int MODEL>  // 0 = trajectory (also writes quaternions), 1 = keyframes (also writes the per-keyframe table)
__global__ void k_pose_chain(PoseBatch pb, float* __restrict__ Mtab, ImuFactors imu, KfFactors kf, TrajTiming tt) {
    DMSA_PDL_ENTER();
    extern __shared__ double sm[];
    const int n = pb.n, v = blockIdx.x, Vld = pb.Vld;
    double* sE = sm;            // exp(relO_k)            [n][9]
    double* sR = sE + 9 * n;    // global rotation R_k    [n][9]
    double* sT = sR + 9 * n;    // global translation     [n][3]
    double* sO = sT + 3 * n;    // global orientation     [n][3]
    const double* p = pb.params + (size_t)v * pb.P;
    // A: exponentials of the relative orientations (parallel over k)
    for (int k = threadIdx.x; k < n; k += blockDim.x) {
        Vec3 w = (k == 0) ? mk3(pb.pose0[0], pb.pose0[1], pb.pose0[2]) : mk3(p[3 * (k - 1)], p[3 * (k - 1) + 1], p[3 * (k - 1) + 2]);
        Mat3 E = so3_exp(w);
#pragma unroll
        for (int i = 0; i < 9; ++i) sE[9 * k + i] = E.m[i];
    }
    __syncthreads();
    // B: the sequential chain (ConsecutivePoses.h:31-42)
    if (threadIdx.x == 0) {
        Mat3 R = identity3();
        Vec3 T = mk3(0, 0, 0);
        const int toff = 3 * (n - 1);
        for (int k = 0; k < n; ++k) {
            Vec3 t = (k == 0) ? mk3(pb.pose0[3], pb.pose0[4], pb.pose0[5]) : mk3(p[toff + 3 * (k - 1)], p[toff + 3 * (k - 1) + 1], p[toff + 3 * (k - 1) + 2]);
            T = add3(T, matvec3(R, t));
            sT[3 * k] = T.x;
            sT[3 * k + 1] = T.y;
            sT[3 * k + 2] = T.z;
            Mat3 E;
#pragma unroll
            for (int i = 0; i < 9; ++i) E.m[i] = sE[9 * k + i];
            R = matmul3(R, E);
#pragma unroll
            for (int i = 0; i < 9; ++i) sR[9 * k + i] = R.m[i];
        }
    }
    __syncthreads();
    // C: logarithms (parallel over k), outputs
    for (int k = threadIdx.x; k < n; k += blockDim.x) {
        Mat3 R;
#pragma unroll
        for (int i = 0; i < 9; ++i) R.m[i] = sR[9 * k + i];
        Vec3 o = so3_log(R);
        sO[3 * k] = o.x;
        sO[3 * k + 1] = o.y;
        sO[3 * k + 2] = o.z;
        pb.globO_t[(size_t)(3 * k + 0) * Vld + v] = o.x;
        pb.globO_t[(size_t)(3 * k + 1) * Vld + v] = o.y;
        pb.globO_t[(size_t)(3 * k + 2) * Vld + v] = o.z;
        pb.globT_t[(size_t)(3 * k + 0) * Vld + v] = sT[3 * k];
        pb.globT_t[(size_t)(3 * k + 1) * Vld + v] = sT[3 * k + 1];
        pb.globT_t[(size_t)(3 * k + 2) * Vld + v] = sT[3 * k + 2];
        if (MODEL == 0) {
            Quat q = quat_from_axang(o);
            pb.quat_t[(size_t)(4 * k + 0) * Vld + v] = q.w;
            pb.quat_t[(size_t)(4 * k + 1) * Vld + v] = q.x;
            pb.quat_t[(size_t)(4 * k + 2) * Vld + v] = q.y;
            pb.quat_t[(size_t)(4 * k + 3) * Vld + v] = q.z;
        } else {
            // MapManagement.h:133-138: currRot = axang2rotm(globalPoses.Orientations.col(k)).cast<float>()
            Mat3 Rr = so3_exp(o);
            float* M = Mtab + ((size_t)k * Vld + v) * 12;
            float4 rows[3];
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                rows[r] = make_float4((float)Rr.m[3 * r], (float)Rr.m[3 * r + 1], (float)Rr.m[3 * r + 2], (float)sT[3 * k + r]);
                reinterpret_cast<float4*>(M)[r] = rows[r];
            }
            store_pair_entry(pb.Mpair, k, v, Vld, rows[0], rows[1], rows[2]);
        }
    }
    if (MODEL == 1 && threadIdx.x == 0) {  // identity row (unused by keyframe points, kept for uniformity)
        float4* M = reinterpret_cast<float4*>(Mtab + ((size_t)n * Vld + v) * 12);
        M[0] = make_float4(1.f, 0.f, 0.f, 0.f);
        M[1] = make_float4(0.f, 1.f, 0.f, 0.f);
        M[2] = make_float4(0.f, 0.f, 1.f, 0.f);
        store_pair_entry(pb.Mpair, n, v, Vld, M[0], M[1], M[2]);
    }
    if (pb.extra == nullptr) return;
    __syncthreads();
    // D: additional residual rows
    if (MODEL == 0) {
        if (!imu.enabled) return;
        // ContinuousTrajectory.h:603-663.  global2relative (:606) re-derives the relative orientations.
        for (int k = 1 + threadIdx.x; k < n; k += blockDim.x) {
            Mat3 Rs = so3_exp(mk3(sO[3 * (k - 1)], sO[3 * (k - 1) + 1], sO[3 * (k - 1) + 2]));
            Mat3 Rk = so3_exp(mk3(sO[3 * k], sO[3 * k + 1], sO[3 * k + 2]));
            Vec3 relO = so3_log(matmul3(transpose3(Rs), Rk));
            double dt = imu.stamps[k] - imu.stamps[k - 1];
            double inv_dt = 1.0 / imu.dt_res;
            int ia = imu.paramIdx[k - 1], ib = imu.paramIdx[k];
            // dense translations at four samples (barycentric-rational interpolation, :214-217)
            Vec3 D[4];
            int js[4] = {ia + 1, ia, ib, ib - 1};
            for (int q = 0; q < 4; ++q) {
                int j = js[q];
                double out[3];
                int hi = tt.hit[j];
                for (int a = 0; a < 3; ++a) {
                    if (hi >= 0) {
                        out[a] = sT[3 * hi + a];
                    } else {
                        double num = 0, den = 0;
                        for (int i = 0; i < n; ++i) {
                            double w = tt.fh[(size_t)j * n + i];
                            num += w * sT[3 * i + a];
                            den += w;
                        }
                        out[a] = num / den;
                    }
                }
                D[q] = mk3(out[0], out[1], out[2]);
            }
            Vec3 vs = scale3(inv_dt, sub3(D[0], D[1]));
            Vec3 ve = scale3(inv_dt, sub3(D[2], D[3]));
            Vec3 grav = mk3(imu.gravity[0], imu.gravity[1], imu.gravity[2]);
            Vec3 Tk = mk3(sT[3 * k], sT[3 * k + 1], sT[3 * k + 2]);
            Vec3 Tp = mk3(sT[3 * (k - 1)], sT[3 * (k - 1) + 1], sT[3 * (k - 1) + 2]);
            Mat3 RsT = transpose3(Rs);
            Vec3 dpm = matvec3(RsT, sub3(sub3(sub3(Tk, Tp), scale3(dt, vs)), scale3(0.5 * dt * dt, grav)));
            Vec3 pos_err = sub3(dpm, mk3(imu.prePos[3 * k], imu.prePos[3 * k + 1], imu.prePos[3 * k + 2]));
            Mat3 Pm;
#pragma unroll
            for (int i = 0; i < 9; ++i) Pm.m[i] = imu.preRot[9 * k + i];
            Vec3 rot_err = so3_log(matmul3(transpose3(Pm), so3_exp(relO)));
            Vec3 dvm = matvec3(RsT, sub3(sub3(ve, vs), scale3(dt, grav)));
            Vec3 vel_err = sub3(dvm, mk3(imu.preVel[3 * k], imu.preVel[3 * k + 1], imu.preVel[3 * k + 2]));
            double r[9] = {rot_err.x, rot_err.y, rot_err.z, vel_err.x, vel_err.y, vel_err.z, pos_err.x, pos_err.y, pos_err.z};
            const double* C = imu.covInv + 81 * (size_t)k;
            double q = 0;
            for (int a = 0; a < 9; ++a) {
                double t = 0;
                for (int b = 0; b < 9; ++b) t += C[9 * a + b] * r[b];
                q += r[a] * t;
            }
            q *= imu.balancing;
            pb.extra[(size_t)(k - 1) * Vld + v] = sqrt(q);
        }
    } else {
        int off = 0;
        if (kf.useGrav) {
            const double ci = 1.0 / (0.3 * 0.3);
            for (int k = threadIdx.x; k < n; k += blockDim.x) {
                double e = 0.0;
                if (k >= 1 && kf.plausible[k]) {
                    Vec3 d = sub3(matvec3(so3_exp(mk3(sO[3 * k], sO[3 * k + 1], sO[3 * k + 2])), mk3(kf.measGrav[3 * k], kf.measGrav[3 * k + 1], kf.measGrav[3 * k + 2])),
                                  mk3(kf.gravity[0], kf.gravity[1], kf.gravity[2]));
                    double q = ci * dot3(d, d);
                    q *= kf.balanceGrav;
                    e = sqrt(q);
                }
                pb.extra[(size_t)k * Vld + v] = e;
            }
            off = n;
        }
        if (kf.useOdom) {
            const double ci = 1.0 / (0.01 * 0.01);
            const int toff = 3 * (n - 1);
            for (int k = 1 + threadIdx.x; k < n; k += blockDim.x) {
                Vec3 relT = mk3(p[toff + 3 * (k - 1)], p[toff + 3 * (k - 1) + 1], p[toff + 3 * (k - 1) + 2]);
                Vec3 relO = mk3(p[3 * (k - 1)], p[3 * (k - 1) + 1], p[3 * (k - 1) + 2]);
                Vec3 td = sub3(mk3(kf.odomT[3 * k], kf.odomT[3 * k + 1], kf.odomT[3 * k + 2]), relT);
                Mat3 Rm;
#pragma unroll
                for (int i = 0; i < 9; ++i) Rm.m[i] = kf.odomR[9 * k + i];
                Vec3 od = so3_log(matmul3(transpose3(so3_exp(relO)), Rm));
                double q = ci * dot3(td, td);
                q += ci * dot3(od, od);
                q *= kf.balanceOdom;
                pb.extra[(size_t)(off + k - 1) * Vld + v] = sqrt(q);
            }
        }
    }
}

// Dense pose of sample j for vector v: orientation (:194-198, :570-591) and translation (:201-218)
__device__ __forceinline__ void dense_pose(const PoseBatch& pb, const TrajTiming& tt, int v, int j, Vec3& aa, double T[3]) {
    const int n = pb.n, Vld = pb.Vld;
    const int r = tt.seg[j];
    if (r > 0) {
        Quat q1, q2;
        q1.w = pb.quat_t[(size_t)(4 * (r - 1) + 0) * Vld + v];
        q1.x = pb.quat_t[(size_t)(4 * (r - 1) + 1) * Vld + v];
        q1.y = pb.quat_t[(size_t)(4 * (r - 1) + 2) * Vld + v];
        q1.z = pb.quat_t[(size_t)(4 * (r - 1) + 3) * Vld + v];
        q2.w = pb.quat_t[(size_t)(4 * r + 0) * Vld + v];
        q2.x = pb.quat_t[(size_t)(4 * r + 1) * Vld + v];
        q2.y = pb.quat_t[(size_t)(4 * r + 2) * Vld + v];
        q2.z = pb.quat_t[(size_t)(4 * r + 3) * Vld + v];
        aa = axang_from_quat(quat_slerp(q1, q2, tt.urel[j]));
    } else {
        aa = mk3(pb.globO_t[(size_t)0 * Vld + v], pb.globO_t[(size_t)1 * Vld + v], pb.globO_t[(size_t)2 * Vld + v]);
    }
    // Boost barycentric_rational::operator()
    const int hi = tt.hit[j];
    if (hi >= 0) {
#pragma unroll
        for (int a = 0; a < 3; ++a) T[a] = pb.globT_t[(size_t)(3 * hi + a) * Vld + v];
    } else {
        double num[3] = {0, 0, 0}, den = 0;
        const double* w = tt.fh + (size_t)j * n;
        for (int i = 0; i < n; ++i) {
            double wi = w[i];
            num[0] += wi * pb.globT_t[(size_t)(3 * i + 0) * Vld + v];
            num[1] += wi * pb.globT_t[(size_t)(3 * i + 1) * Vld + v];
            num[2] += wi * pb.globT_t[(size_t)(3 * i + 2) * Vld + v];
            den += wi;
        }
#pragma unroll
        for (int a = 0; a < 3; ++a) T[a] = num[a] / den;
    }
}

// Dense trajectory table: thread (v, j).  ContinuousTrajectory.h:194-225.
__global__ void k_dense_table(PoseBatch pb, TrajTiming tt, float* __restrict__ Mtab) {
    DMSA_PDL_ENTER();
    const int v = blockIdx.x * 32 + threadIdx.x;
    const int j = blockIdx.y * blockDim.y + threadIdx.y;
    if (v >= pb.V || j > tt.n_total) return;
    const int Vld = pb.Vld;
    if (j == tt.n_total) {  // identity row: static points (already in the world frame) go through the same code path
        float4* M = reinterpret_cast<float4*>(Mtab + ((size_t)j * Vld + v) * 12);
        M[0] = make_float4(1.f, 0.f, 0.f, 0.f);
        M[1] = make_float4(0.f, 1.f, 0.f, 0.f);
        M[2] = make_float4(0.f, 0.f, 1.f, 0.f);
        store_pair_entry(pb.Mpair, j, v, Vld, M[0], M[1], M[2]);
        return;
    }
    Vec3 aa;
    double T[3];
    dense_pose(pb, tt, v, j, aa, T);
    // :221-225
    Mat3 R = so3_exp(aa);
    float4* M = reinterpret_cast<float4*>(Mtab + ((size_t)j * Vld + v) * 12);
    const float4 r0 = make_float4((float)R.m[0], (float)R.m[1], (float)R.m[2], (float)T[0]);
    const float4 r1 = make_float4((float)R.m[3], (float)R.m[4], (float)R.m[5], (float)T[1]);
    const float4 r2 = make_float4((float)R.m[6], (float)R.m[7], (float)R.m[8], (float)T[2]);
    M[0] = r0;
    M[1] = r1;
    M[2] = r2;
    store_pair_entry(pb.Mpair, j, v, Vld, r0, r1, r2);
}

// denseGlobalPoses of vector v (the double poses behind the table): orient / transl are 3 x n_total column-major
__global__ void k_dense_poses(PoseBatch pb, TrajTiming tt, int v, double* __restrict__ orient, double* __restrict__ transl) {
    DMSA_PDL_ENTER();
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= tt.n_total) return;
    Vec3 aa;
    double T[3];
    dense_pose(pb, tt, v, j, aa, T);
    orient[3 * (size_t)j] = aa.x;
    orient[3 * (size_t)j + 1] = aa.y;
    orient[3 * (size_t)j + 2] = aa.z;
    transl[3 * (size_t)j] = T[0];
    transl[3 * (size_t)j + 1] = T[1];
    transl[3 * (size_t)j + 2] = T[2];
}

}  // namespace dmsa
