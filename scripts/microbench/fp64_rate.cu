// fp64_rate.cu — throughput (lane-ops / clk / SM) and dependent latency of DADD, DMUL, F2F.F64.F32, F2F.F32.F64, I2F, IADD3 (64-bit add)
// on the device at hand.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false -O3 -o fp64_rate fp64_rate.cu
#include <cuda_runtime.h>
#include <cstdio>
#define ITER 4096
template <int OP>
__global__ void __launch_bounds__(1024) k_tput(double* out, float fseed, double dseed) {
    double a[8];
    float f[8];
    long long q[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        a[u] = dseed + u + threadIdx.x;
        f[u] = fseed + u + threadIdx.x;
        q[u] = (long long)(u + threadIdx.x);
    }
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (OP == 0) a[u] = a[u] + dseed;                    // DADD
            if (OP == 1) a[u] = a[u] * dseed;                    // DMUL
            if (OP == 2) { a[u] = (double)f[u]; f[u] = f[u] + 1.0f; }   // F2F.F64.F32 (+ 1 FADD)
            if (OP == 3) { f[u] = (float)a[u]; a[u] = __longlong_as_double(__double_as_longlong(a[u]) + 1); }  // F2F.F32.F64 (+ int add)
            if (OP == 4) q[u] = q[u] + (long long)it * 3 + u;    // 64-bit integer add
            if (OP == 5) f[u] = f[u] + fseed;                    // FADD (reference)
        }
    }
    double s = 0;
#pragma unroll
    for (int u = 0; u < 8; ++u) s += a[u] + (double)f[u] + (double)q[u];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int OP>
__global__ void k_lat(double* out, long long* cyc, double dseed, float fseed) {
    double a = dseed;
    float f = fseed;
    long long t0 = clock64();
    for (int it = 0; it < ITER; ++it) {
        if (OP == 0) a = a + dseed;
        if (OP == 1) a = a * dseed;
        if (OP == 2) { a = (double)f; f = (float)a + 1.0f; }
        if (OP == 5) f = f + fseed;
    }
    long long t1 = clock64();
    out[0] = a + f;
    cyc[0] = t1 - t0;
}
int main() {
    double* out;
    long long* cyc;
    cudaMalloc(&out, 148 * 2 * 1024 * 8);
    cudaMalloc(&cyc, 8);
    cudaDeviceProp pr;
    cudaGetDeviceProperties(&pr, 0);
    const double ghz = 1.965;
    const char* names[6] = {"DADD", "DMUL", "F2F.F64.F32(+FADD)", "F2F.F32.F64(+IADD)", "IADD64", "FADD"};
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int op = 0; op < 6; ++op) {
        float best = 1e9;
        for (int rep = 0; rep < 4; ++rep) {
            cudaEventRecord(e0);
            switch (op) {
                case 0: k_tput<0><<<148 * 2, 1024>>>(out, 1.f, 1.0000001); break;
                case 1: k_tput<1><<<148 * 2, 1024>>>(out, 1.f, 1.0000001); break;
                case 2: k_tput<2><<<148 * 2, 1024>>>(out, 1.f, 1.0000001); break;
                case 3: k_tput<3><<<148 * 2, 1024>>>(out, 1.f, 1.0000001); break;
                case 4: k_tput<4><<<148 * 2, 1024>>>(out, 1.f, 1.0000001); break;
                case 5: k_tput<5><<<148 * 2, 1024>>>(out, 1.f, 1.0000001); break;
            }
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            if (rep && ms < best) best = ms;
        }
        const double ops = 148.0 * 2 * 1024 * ITER * 8;
        printf("%-22s %8.4f ms  %7.2f lane-ops/clk/SM (at %.3f GHz, %d SMs)\n", names[op], best, ops / (best * 1e-3) / (ghz * 1e9) / pr.multiProcessorCount, ghz,
               pr.multiProcessorCount);
    }
    long long h;
    k_lat<0><<<1, 1>>>(out, cyc, 1.0000001, 1.f); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); printf("DADD dependent latency %.1f cycles\n", (double)h / ITER);
    k_lat<1><<<1, 1>>>(out, cyc, 1.0000001, 1.f); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); printf("DMUL dependent latency %.1f cycles\n", (double)h / ITER);
    k_lat<2><<<1, 1>>>(out, cyc, 1.0000001, 1.f); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); printf("F2F64+F2F32+FADD round trip %.1f cycles\n", (double)h / ITER);
    k_lat<5><<<1, 1>>>(out, cyc, 1.0000001, 1.f); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); printf("FADD dependent latency %.1f cycles\n", (double)h / ITER);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
