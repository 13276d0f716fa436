// FP64 pipe microbenchmark: can 52-bit limb products (two round-toward-zero FMAs + one add, Emmart et al.)
// beat the half-rate IMAD.WIDE.U32 on B200?  All operands are loop-carried so nothing can be hoisted.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
constexpr int ITERS = 4096;
template <int MODE> __global__ void k(double* out, double seed) {
    double x[8], y = seed + threadIdx.x * 1e-3;
    unsigned long long acc[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { x[i] = seed * (i + 1) + threadIdx.x; acc[i] = i; }
    const double c1 = 1329227995784915872903807060280344576.0;  // 2^120 (any large constant)
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (MODE == 0) x[i] = __fma_rz(x[i], y, x[(i + 4) & 7]);                 // plain dependent DFMA chains
            if (MODE == 1) {                                                         // hi/lo product + integer accumulate
                double hi = __fma_rz(x[i], y, c1);
                double sub = c1 - hi;
                double lo = __fma_rz(x[i], y, sub);
                acc[i] += (unsigned long long)__double_as_longlong(hi);
                acc[(i + 1) & 7] += (unsigned long long)__double_as_longlong(lo);
                x[i] = lo * 0.5 + 1.0;                                               // keep operands changing (extra DFMA)
            }
            if (MODE == 2) { x[i] = x[i] + y; }                                      // DADD
        }
    }
    double s = 0; unsigned long long t = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) { s += x[i]; t += acc[i]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + (double)t;
}
template <class K> float time_ms(K launch) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    launch(); launch(); cudaDeviceSynchronize();
    cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main() {
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    int sms = prop.multiProcessorCount;
    double* out; cudaMalloc(&out, (size_t)sms * 2 * 1024 * 8);
    for (int warps = 2; warps <= 32; warps *= 2) {
        int threads = warps * 32, blocks = sms * 2;
        double n = (double)blocks * threads * ITERS * 8;
        float t0 = time_ms([&] { k<0><<<blocks, threads>>>(out, 1.5); });
        float t1 = time_ms([&] { k<1><<<blocks, threads>>>(out, 1.5); });
        float t2 = time_ms([&] { k<2><<<blocks, threads>>>(out, 1.5); });
        printf("{\"warps_per_sm\": %d, \"dfma_T_per_s\": %.3f, \"dfma_lanes_per_clk_per_sm\": %.1f, \"hilo_products_T_per_s\": %.3f, \"hilo_fp64ops_lanes_per_clk_per_sm\": %.1f, \"dadd_T_per_s\": %.3f}\n",
               warps * 2, n / t0 / 1e9, n / t0 / 1e3 / sms / 1.965e6, n / t1 / 1e9, 4 * n / t1 / 1e3 / sms / 1.965e6, n / t2 / 1e9);
    }
    return 0;
}
