// se3_math.cuh — double-precision SO(3)/SE(3) helpers shared by host and device code of the
// product (NOT shared with oracle/: the oracle keeps its own independent restatement).
//
// Restates, for the GPU path:
//   helpers.h:24-37   slerp (AngleAxisd -> Quaterniond -> Quaterniond::slerp -> AngleAxisd)
//   helpers.h:51-57   axang2rotm (identity below EPSILON_ROT = 1e-5, else matrix exponential)
//   helpers.h:59-65   rotm2axang (vee of the principal matrix logarithm)
//   ConsecutivePoses.h:26-67 relative2global / global2relative
#pragma once
#include <cmath>
#include <cstdint>

#ifdef __CUDACC__
#define DMSA_HD __host__ __device__ __forceinline__
#else
#define DMSA_HD inline
#endif

namespace dmsa {

struct Vec3 {
    double x, y, z;
};
struct Mat3 {
    double m[9];  // row-major
};
struct Quat {
    double w, x, y, z;
};

DMSA_HD Vec3 mk3(double x, double y, double z) {
    Vec3 r;
    r.x = x;
    r.y = y;
    r.z = z;
    return r;
}
DMSA_HD double dot3(const Vec3& a, const Vec3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
DMSA_HD double norm3(const Vec3& a) { return sqrt(dot3(a, a)); }
DMSA_HD Vec3 add3(const Vec3& a, const Vec3& b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
DMSA_HD Vec3 sub3(const Vec3& a, const Vec3& b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
DMSA_HD Vec3 scale3(double s, const Vec3& a) { return mk3(s * a.x, s * a.y, s * a.z); }

DMSA_HD Mat3 identity3() {
    Mat3 r;
    r.m[0] = 1; r.m[1] = 0; r.m[2] = 0;
    r.m[3] = 0; r.m[4] = 1; r.m[5] = 0;
    r.m[6] = 0; r.m[7] = 0; r.m[8] = 1;
    return r;
}
DMSA_HD Mat3 matmul3(const Mat3& a, const Mat3& b) {
    Mat3 r;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) r.m[i * 3 + j] = a.m[i * 3] * b.m[j] + a.m[i * 3 + 1] * b.m[3 + j] + a.m[i * 3 + 2] * b.m[6 + j];
    return r;
}
DMSA_HD Mat3 transpose3(const Mat3& a) {
    Mat3 r;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) r.m[i * 3 + j] = a.m[j * 3 + i];
    return r;
}
DMSA_HD Vec3 matvec3(const Mat3& a, const Vec3& v) {
    return mk3(a.m[0] * v.x + a.m[1] * v.y + a.m[2] * v.z, a.m[3] * v.x + a.m[4] * v.y + a.m[5] * v.z, a.m[6] * v.x + a.m[7] * v.y + a.m[8] * v.z);
}

// helpers.h:51-57: Rodrigues form of exp(skew(w)); identity for ||w|| < 1e-5.
DMSA_HD Mat3 so3_exp(const Vec3& w) {
    double th = norm3(w);
    if (th < 0.00001) return identity3();
    double s, c;
#ifdef __CUDA_ARCH__
    sincos(th, &s, &c);
#else
    s = sin(th);
    c = cos(th);
#endif
    double a = s / th, b = (1.0 - c) / (th * th);
    double x = w.x, y = w.y, z = w.z;
    Mat3 r;
    r.m[0] = 1.0 + b * (-(y * y) - z * z);
    r.m[1] = -a * z + b * (x * y);
    r.m[2] = a * y + b * (x * z);
    r.m[3] = a * z + b * (x * y);
    r.m[4] = 1.0 + b * (-(x * x) - z * z);
    r.m[5] = -a * x + b * (y * z);
    r.m[6] = -a * y + b * (x * z);
    r.m[7] = a * x + b * (y * z);
    r.m[8] = 1.0 + b * (-(x * x) - y * y);
    return r;
}

// helpers.h:59-65: principal logarithm.
DMSA_HD Vec3 so3_log(const Mat3& R) {
    Vec3 v = mk3(0.5 * (R.m[7] - R.m[5]), 0.5 * (R.m[2] - R.m[6]), 0.5 * (R.m[3] - R.m[1]));
    double s = norm3(v);
    double c = 0.5 * (R.m[0] + R.m[4] + R.m[8] - 1.0);
    double th = atan2(s, c);
    if (s > 1e-7) return scale3(th / s, v);
    if (c > 0.0) return v;
    // rotation by ~pi: axis from the symmetric part R ~ 2 a a^T - I
    double ax = sqrt(fmax(0.0, 0.5 * (R.m[0] + 1.0)));
    double ay = sqrt(fmax(0.0, 0.5 * (R.m[4] + 1.0)));
    double az = sqrt(fmax(0.0, 0.5 * (R.m[8] + 1.0)));
    if (ax >= ay && ax >= az) {
        if (R.m[1] + R.m[3] < 0) ay = -ay;
        if (R.m[2] + R.m[6] < 0) az = -az;
    } else if (ay >= az) {
        if (R.m[1] + R.m[3] < 0) ax = -ax;
        if (R.m[5] + R.m[7] < 0) az = -az;
    } else {
        if (R.m[2] + R.m[6] < 0) ax = -ax;
        if (R.m[5] + R.m[7] < 0) ay = -ay;
    }
    Vec3 a = mk3(ax, ay, az);
    double n = norm3(a);
    if (n > 0) a = scale3(1.0 / n, a);
    if (dot3(a, v) < 0) a = scale3(-1.0, a);
    return scale3(th, a);
}

// helpers.h:26: Quaterniond(AngleAxisd(aa.norm(), aa.normalized()))
DMSA_HD Quat quat_from_axang(const Vec3& aa) {
    double z = dot3(aa, aa);
    double ang = sqrt(z);
    Vec3 axis = aa;
    if (z > 0.0) axis = scale3(1.0 / sqrt(z), aa);
    double s, c;
#ifdef __CUDA_ARCH__
    sincos(0.5 * ang, &s, &c);
#else
    s = sin(0.5 * ang);
    c = cos(0.5 * ang);
#endif
    Quat q;
    q.w = c;
    q.x = s * axis.x;
    q.y = s * axis.y;
    q.z = s * axis.z;
    return q;
}

// Eigen 3.4 QuaternionBase::slerp(t, other)
DMSA_HD Quat quat_slerp(const Quat& a, const Quat& b, double t) {
    const double one = 1.0 - 2.220446049250313e-16;
    double d = a.w * b.w + a.x * b.x + a.y * b.y + a.z * b.z;
    double absD = fabs(d);
    double s0, s1;
    if (absD >= one) {
        s0 = 1.0 - t;
        s1 = t;
    } else {
        double theta = acos(absD);
        double sinTheta = sin(theta);
        s0 = sin((1.0 - t) * theta) / sinTheta;
        s1 = sin(t * theta) / sinTheta;
    }
    if (d < 0.0) s1 = -s1;
    Quat q;
    q.w = s0 * a.w + s1 * b.w;
    q.x = s0 * a.x + s1 * b.x;
    q.y = s0 * a.y + s1 * b.y;
    q.z = s0 * a.z + s1 * b.z;
    return q;
}

// Eigen 3.4 AngleAxis(const QuaternionBase&) then helpers.h:33 axis * angle
DMSA_HD Vec3 axang_from_quat(const Quat& q) {
    double n = sqrt(q.x * q.x + q.y * q.y + q.z * q.z);
    if (n != 0.0) {
        double angle = 2.0 * atan2(n, fabs(q.w));
        if (q.w < 0.0) n = -n;
        return mk3(q.x / n * angle, q.y / n * angle, q.z / n * angle);
    }
    return mk3(0.0, 0.0, 0.0);
}

}  // namespace dmsa
