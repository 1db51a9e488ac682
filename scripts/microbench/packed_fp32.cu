// packed_fp32.cu — micro-benchmark: does Blackwell's packed FP32x2 (FMUL2/FADD2) relieve the issue-bound cost loop
// WITHOUT fusing multiplies into adds?  ptxas 12.9 contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 even under
// --fmad=false, so the bit-faithful variant keeps every multiply->add edge with one scalar side:
//   products packed (FMUL2 on vector pairs), the adds they feed scalar (FADD), add-after-add packed (FADD2).
// Variants: 0 = scalar reference loop (production arithmetic), 1 = mixed packed (bit-identical by construction),
//           2 = everything packed, fusion allowed (NOT faithful; upper bound of the pipe).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false -O3 -lineinfo -o packed_fp32 packed_fp32.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float lo, float hi) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void upk(u64 p, float& lo, float& hi) { asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(p)); }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 sub2(u64 a, u64 b) { u64 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 mul2s(u64 a, float s) { return mul2(a, pk(s, s)); }

#define NMEM 256
struct Rows { u64 m[12]; };  // pair-interleaved 3x4 transform of vectors (2t, 2t+1)

// ---- variant 0: scalar, one vector per thread --------------------------------------------------------------
__global__ void __launch_bounds__(128, 10) k_scalar(const float4* __restrict__ rec, const float4* __restrict__ Mtab, int Vld, const float* __restrict__ info,
                                                    double* __restrict__ out, int nrows) {
    __shared__ float4 s[NMEM];
    for (int i = threadIdx.x; i < NMEM; i += blockDim.x) s[i] = rec[(size_t)blockIdx.x * NMEM + i];
    __syncthreads();
    const int v = threadIdx.x;
    double sx = 0, sy = 0, sz = 0;
    int tprev = -1;
    float4 m0, m1, m2;
    m0 = m1 = m2 = make_float4(0, 0, 0, 0);
#define ROWUP(t)                                                      \
    if (t != tprev) {                                                 \
        const float4* Mp = Mtab + ((size_t)t * Vld + v) * 3;          \
        m0 = __ldg(Mp); m1 = __ldg(Mp + 1); m2 = __ldg(Mp + 2);       \
        tprev = t;                                                    \
    }
#define XF(r, X, Y, Z)                                                                                                          \
    X = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(m0.x, r.x), __fmul_rn(m0.y, r.y)), __fmul_rn(m0.z, r.z)), m0.w);                \
    Y = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(m1.x, r.x), __fmul_rn(m1.y, r.y)), __fmul_rn(m1.z, r.z)), m1.w);                \
    Z = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(m2.x, r.x), __fmul_rn(m2.y, r.y)), __fmul_rn(m2.z, r.z)), m2.w);
#pragma unroll 4
    for (int j = 0; j < NMEM; ++j) {
        const float4 r = s[j];
        const int t = __float_as_int(r.w);
        ROWUP(t)
        float X, Y, Z;
        XF(r, X, Y, Z)
        sx += (double)X; sy += (double)Y; sz += (double)Z;
    }
    const float nf = (float)NMEM;
    const float mx = __fdiv_rn((float)sx, nf), my = __fdiv_rn((float)sy, nf), mz = __fdiv_rn((float)sz, nf);
    const float* I = info + 9 * blockIdx.x;
    const float i0 = I[0], i1 = I[1], i2 = I[2], i3 = I[3], i4 = I[4], i5 = I[5], i6 = I[6], i7 = I[7], i8 = I[8];
    const float wk = 0.37f;
    double acc = 0;
    tprev = -1;
#pragma unroll 4
    for (int j = 0; j < NMEM; ++j) {
        const float4 r = s[j];
        const int t = __float_as_int(r.w);
        ROWUP(t)
        float X, Y, Z;
        XF(r, X, Y, Z)
        const float d0 = __fsub_rn(X, mx), d1 = __fsub_rn(Y, my), d2 = __fsub_rn(Z, mz);
        const float t0 = __fmul_rn(wk, d0), t1 = __fmul_rn(wk, d1), t2 = __fmul_rn(wk, d2);
        const float r0 = __fadd_rn(__fmul_rn(t0, i0), __fadd_rn(__fmul_rn(t1, i3), __fmul_rn(t2, i6)));
        const float r1 = __fadd_rn(__fmul_rn(t0, i1), __fadd_rn(__fmul_rn(t1, i4), __fmul_rn(t2, i7)));
        const float r2 = __fadd_rn(__fmul_rn(t0, i2), __fadd_rn(__fmul_rn(t1, i5), __fmul_rn(t2, i8)));
        const float s_ = __fadd_rn(__fmul_rn(r0, d0), __fadd_rn(__fmul_rn(r1, d1), __fmul_rn(r2, d2)));
        acc += (double)s_;
    }
    out[(size_t)blockIdx.x * Vld + v] = sqrt(fabs(acc));
}

// ---- variants 1 / 2: two vectors per thread, pair-interleaved table --------------------------------------------
// Mpair[((row * Vp + tp) * 12 + c)] = {M[row][2tp][c], M[row][2tp+1][c]} as one 64-bit word, c = r*4 + col
template <int FUSE_OK>
__global__ void __launch_bounds__(64, 16) k_packed(const float4* __restrict__ rec, const u64* __restrict__ Mpair, int Vp, const float* __restrict__ info,
                                                   double* __restrict__ out, int Vld) {
    __shared__ float4 s[NMEM];
    for (int i = threadIdx.x; i < NMEM; i += blockDim.x) s[i] = rec[(size_t)blockIdx.x * NMEM + i];
    __syncthreads();
    const int tp = threadIdx.x;
    double sx0 = 0, sy0 = 0, sz0 = 0, sx1 = 0, sy1 = 0, sz1 = 0;
    int tprev = -1;
    u64 m[12];
#pragma unroll
    for (int c = 0; c < 12; ++c) m[c] = 0;
#define ROWUP2(t)                                                                             \
    if (t != tprev) {                                                                         \
        const ulonglong2* Mp = reinterpret_cast<const ulonglong2*>(Mpair + ((size_t)t * Vp + tp) * 12); \
        _Pragma("unroll") for (int c = 0; c < 6; ++c) {                                       \
            const ulonglong2 q = __ldg(Mp + c);                                               \
            m[2 * c] = q.x; m[2 * c + 1] = q.y;                                               \
        }                                                                                     \
        tprev = t;                                                                            \
    }
    // one row of the transform for the vector pair: products packed, the two adds they feed scalar, "+ m.w" packed
#define XROW(b, r, OUT)                                                                       \
    {                                                                                         \
        const u64 p0 = mul2s(m[b + 0], r.x), p1 = mul2s(m[b + 1], r.y), p2 = mul2s(m[b + 2], r.z); \
        if (FUSE_OK) {                                                                        \
            OUT = add2(add2(add2(p0, p1), p2), m[b + 3]);                                     \
        } else {                                                                              \
            float p0a, p0b, p1a, p1b, p2a, p2b;                                               \
            upk(p0, p0a, p0b); upk(p1, p1a, p1b); upk(p2, p2a, p2b);                          \
            const float qa = __fadd_rn(__fadd_rn(p0a, p1a), p2a);                             \
            const float qb = __fadd_rn(__fadd_rn(p0b, p1b), p2b);                             \
            OUT = add2(pk(qa, qb), m[b + 3]);                                                 \
        }                                                                                     \
    }
#pragma unroll 4
    for (int j = 0; j < NMEM; ++j) {
        const float4 r = s[j];
        const int t = __float_as_int(r.w);
        ROWUP2(t)
        u64 X, Y, Z;
        XROW(0, r, X) XROW(4, r, Y) XROW(8, r, Z)
        float a, b;
        upk(X, a, b); sx0 += (double)a; sx1 += (double)b;
        upk(Y, a, b); sy0 += (double)a; sy1 += (double)b;
        upk(Z, a, b); sz0 += (double)a; sz1 += (double)b;
    }
    const float nf = (float)NMEM;
    const u64 MX = pk(__fdiv_rn((float)sx0, nf), __fdiv_rn((float)sx1, nf));
    const u64 MY = pk(__fdiv_rn((float)sy0, nf), __fdiv_rn((float)sy1, nf));
    const u64 MZ = pk(__fdiv_rn((float)sz0, nf), __fdiv_rn((float)sz1, nf));
    const float* I = info + 9 * blockIdx.x;
    const float i0 = I[0], i1 = I[1], i2 = I[2], i3 = I[3], i4 = I[4], i5 = I[5], i6 = I[6], i7 = I[7], i8 = I[8];
    const float wk = 0.37f;
    double acc0 = 0, acc1 = 0;
    tprev = -1;
    // r = t0*ia + (t1*ib + t2*ic): three packed products, two scalar adds per vector
#define DOT3(P0, P1, P2, OUTA, OUTB)                                                          \
    {                                                                                         \
        if (FUSE_OK) {                                                                        \
            upk(add2(P0, add2(P1, P2)), OUTA, OUTB);                                          \
        } else {                                                                              \
            float a0, b0, a1, b1, a2, b2;                                                     \
            upk(P0, a0, b0); upk(P1, a1, b1); upk(P2, a2, b2);                                \
            OUTA = __fadd_rn(a0, __fadd_rn(a1, a2));                                          \
            OUTB = __fadd_rn(b0, __fadd_rn(b1, b2));                                          \
        }                                                                                     \
    }
#pragma unroll 4
    for (int j = 0; j < NMEM; ++j) {
        const float4 r = s[j];
        const int t = __float_as_int(r.w);
        ROWUP2(t)
        u64 X, Y, Z;
        XROW(0, r, X) XROW(4, r, Y) XROW(8, r, Z)
        const u64 d0 = sub2(X, MX), d1 = sub2(Y, MY), d2 = sub2(Z, MZ);
        const u64 t0 = mul2s(d0, wk), t1 = mul2s(d1, wk), t2 = mul2s(d2, wk);
        float r0a, r0b, r1a, r1b, r2a, r2b, sa, sb;
        DOT3(mul2s(t0, i0), mul2s(t1, i3), mul2s(t2, i6), r0a, r0b)
        DOT3(mul2s(t0, i1), mul2s(t1, i4), mul2s(t2, i7), r1a, r1b)
        DOT3(mul2s(t0, i2), mul2s(t1, i5), mul2s(t2, i8), r2a, r2b)
        DOT3(mul2(pk(r0a, r0b), d0), mul2(pk(r1a, r1b), d1), mul2(pk(r2a, r2b), d2), sa, sb)
        acc0 += (double)sa;
        acc1 += (double)sb;
    }
    out[(size_t)blockIdx.x * Vld + 2 * tp] = sqrt(fabs(acc0));
    out[(size_t)blockIdx.x * Vld + 2 * tp + 1] = sqrt(fabs(acc1));
}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)

int main(int argc, char** argv) {
    const int G = argc > 1 ? atoi(argv[1]) : 148 * 40;  // sets (blocks)
    const int V = 128, Vld = 128, Vp = 64, ROWS = 1001;
    const int RL = argc > 2 ? atoi(argv[2]) : 6;
    std::vector<float4> rec((size_t)G * NMEM);
    std::vector<float> M((size_t)ROWS * Vld * 12), info((size_t)G * 9);
    srand(7);
    auto rnd = [] { return (float)rand() / RAND_MAX * 2.f - 1.f; };
    for (size_t i = 0; i < rec.size(); ++i) {
        int t = (int)((i / RL) % ROWS);  // the row changes every RL-th member
        rec[i] = make_float4(20.f * rnd(), 20.f * rnd(), 3.f * rnd(), 0.f);
        memcpy(&rec[i].w, &t, 4);
    }
    for (auto& x : M) x = rnd();
    for (auto& x : info) x = 10.f * rnd();
    std::vector<u64> Mp((size_t)ROWS * Vp * 12);
    for (int row = 0; row < ROWS; ++row)
        for (int tp = 0; tp < Vp; ++tp)
            for (int c = 0; c < 12; ++c) {
                float lo = M[((size_t)row * Vld + 2 * tp) * 12 + c], hi = M[((size_t)row * Vld + 2 * tp + 1) * 12 + c];
                unsigned a, b;
                memcpy(&a, &lo, 4);
                memcpy(&b, &hi, 4);
                Mp[((size_t)row * Vp + tp) * 12 + c] = (u64)a | ((u64)b << 32);
            }
    float4 *d_rec, *d_M;
    u64* d_Mp;
    float* d_info;
    double* d_out[3];
    CK(cudaMalloc(&d_rec, rec.size() * 16));
    CK(cudaMalloc(&d_M, M.size() * 4));
    CK(cudaMalloc(&d_Mp, Mp.size() * 8));
    CK(cudaMalloc(&d_info, info.size() * 4));
    for (int k = 0; k < 3; ++k) CK(cudaMalloc(&d_out[k], (size_t)G * Vld * 8));
    CK(cudaMemcpy(d_rec, rec.data(), rec.size() * 16, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_M, M.data(), M.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_Mp, Mp.data(), Mp.size() * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_info, info.data(), info.size() * 4, cudaMemcpyHostToDevice));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float ms[3] = {0, 0, 0};
    for (int rep = 0; rep < 6; ++rep) {
        for (int var = 0; var < 3; ++var) {
            cudaEventRecord(e0);
            if (var == 0) k_scalar<<<G, V>>>(d_rec, d_M, Vld, d_info, d_out[0], ROWS);
            if (var == 1) k_packed<0><<<G, Vp>>>(d_rec, d_Mp, Vp, d_info, d_out[1], Vld);
            if (var == 2) k_packed<1><<<G, Vp>>>(d_rec, d_Mp, Vp, d_info, d_out[2], Vld);
            cudaEventRecord(e1);
            CK(cudaEventSynchronize(e1));
            float t;
            cudaEventElapsedTime(&t, e0, e1);
            if (rep >= 1) ms[var] = (rep == 1) ? t : (t < ms[var] ? t : ms[var]);
        }
    }
    CK(cudaGetLastError());
    std::vector<double> o0((size_t)G * Vld), o1(o0.size()), o2(o0.size());
    CK(cudaMemcpy(o0.data(), d_out[0], o0.size() * 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(o1.data(), d_out[1], o0.size() * 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(o2.data(), d_out[2], o0.size() * 8, cudaMemcpyDeviceToHost));
    size_t diff1 = 0, diff2 = 0;
    for (size_t i = 0; i < o0.size(); ++i) {
        diff1 += memcmp(&o0[i], &o1[i], 8) != 0;
        diff2 += memcmp(&o0[i], &o2[i], 8) != 0;
    }
    const double mv = (double)G * NMEM * V;
    printf("{\"run_length\": %d, \"member_vectors\": %.0f, \"ms_scalar\": %.4f, \"ms_mixed_packed\": %.4f, \"ms_all_packed_fused\": %.4f, "
           "\"mv_per_ns_scalar\": %.2f, \"mv_per_ns_mixed\": %.2f, \"mv_per_ns_fused\": %.2f, "
           "\"bitdiff_mixed_vs_scalar\": %zu, \"bitdiff_fused_vs_scalar\": %zu, \"of\": %zu}\n",
           RL, mv, ms[0], ms[1], ms[2], mv / ms[0] * 1e-6, mv / ms[1] * 1e-6, mv / ms[2] * 1e-6, diff1, diff2, o0.size());
    return 0;
}
