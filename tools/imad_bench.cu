// Integer-pipe microbenchmark for the roofline denominator (IMAD peak is not in
// MEASURED_PEAKS.json).  Register-only loops, no memory traffic.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/imad_bench tools/imad_bench.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include "../bgls_b200/csrc/field.cuh"
using namespace bgls;

constexpr int ITERS = 4096;

template <int MODE> __global__ void k_imad(uint32_t* out, uint32_t seed) {
    uint32_t a = seed + threadIdx.x, b = seed * 3 + 1;
    uint32_t x[8];
    uint64_t w[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { x[i] = a + i; w[i] = a * 7 + i; }
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (MODE == 0) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(b), "r"(a));
            if (MODE == 1) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(b), "r"(a));
            // multiplier = low word of a neighbouring accumulator: a loop-invariant operand lets ptxas hoist the product
            if (MODE == 2) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"((uint32_t)w[(i + 1) & 7]), "r"(b));
            if (MODE == 6) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(x[i]), "r"(b));
        }
        if (MODE == 3) {  // one carry chain of 8 fused IMAD.WIDE.X per iteration
            uint32_t* p = (uint32_t*)w;
            asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(p[0]), "+r"(p[1]) : "r"(a), "r"(b));
#pragma unroll
            for (int i = 1; i < 8; i++)
                asm volatile("madc.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(p[2 * i]), "+r"(p[2 * i + 1]) : "r"(a), "r"(b));
        }
        if (MODE == 4) {  // two independent carry chains of 4
            uint32_t* p = (uint32_t*)w;
            asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(p[0]), "+r"(p[1]) : "r"(a), "r"(b));
#pragma unroll
            for (int i = 1; i < 4; i++)
                asm volatile("madc.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(p[2 * i]), "+r"(p[2 * i + 1]) : "r"(a), "r"(b));
            asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(p[8]), "+r"(p[9]) : "r"(b), "r"(a));
#pragma unroll
            for (int i = 5; i < 8; i++)
                asm volatile("madc.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(p[2 * i]), "+r"(p[2 * i + 1]) : "r"(b), "r"(a));
        }
        if (MODE == 5) {  // IADD3 chain
#pragma unroll
            for (int i = 0; i < 8; i++) asm volatile("add.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(b));
        }
    }
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += x[i] + (uint32_t)w[i] + (uint32_t)(w[i] >> 32);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class C> __global__ void k_fpmul(uint32_t* out, uint32_t seed, int iters) {
    Fp<C> a, b;
#pragma unroll
    for (int i = 0; i < C::N; i++) { a.v[i] = seed + threadIdx.x + i; b.v[i] = seed * 5 + i; }
    a.v[C::N - 1] &= 0x0fffffff; b.v[C::N - 1] &= 0x0fffffff;
    for (int it = 0; it < iters; it++) { fp_mul(a, a, b); fp_mul(b, b, a); }
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < C::N; i++) s += a.v[i] ^ b.v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class K> float time_ms(K launch) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    launch(); launch();
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    launch();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}

int main() {
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    int sms = prop.multiProcessorCount;
    uint32_t* out; cudaMalloc(&out, 148 * 64 * 1024 * 4);
    const char* names[7] = {"imad_lo", "imad_hi", "imad_wide", "imad_wide_carry_chain8", "imad_wide_carry_2chains4", "iadd", "imad_wide_invariant_operand(hoisted:64-bit adds)"};
    printf("{\"device\": \"%s\", \"sms\": %d, \"clock_khz\": %d}\n", prop.name, sms, prop.clockRate);
    for (int warps = 2; warps <= 32; warps *= 2) {
        int threads = warps * 32, blocks = sms * 2;
        double ops = (double)blocks * threads * ITERS * 8;
        float t[7];
        t[0] = time_ms([&] { k_imad<0><<<blocks, threads>>>(out, 3); });
        t[1] = time_ms([&] { k_imad<1><<<blocks, threads>>>(out, 3); });
        t[2] = time_ms([&] { k_imad<2><<<blocks, threads>>>(out, 3); });
        t[3] = time_ms([&] { k_imad<3><<<blocks, threads>>>(out, 3); });
        t[4] = time_ms([&] { k_imad<4><<<blocks, threads>>>(out, 3); });
        t[5] = time_ms([&] { k_imad<5><<<blocks, threads>>>(out, 3); });
        t[6] = time_ms([&] { k_imad<6><<<blocks, threads>>>(out, 3); });
        for (int m = 0; m < 7; m++)
            printf("{\"bench\": \"%s\", \"warps_per_sm\": %d, \"Tops_per_s\": %.3f, \"ops_per_clk_per_sm_at_1965MHz\": %.1f}\n", names[m], warps * 2,
                   ops / t[m] / 1e9, ops / t[m] / 1e3 / sms / 1.965e6);
    }
    for (int warps = 1; warps <= 16; warps *= 2) {
        int threads = warps * 32, blocks = sms * 2, iters = 2048;
        double muls = (double)blocks * threads * iters * 2;
        float t8 = time_ms([&] { k_fpmul<BN254><<<blocks, threads>>>(out, 3, iters); });
        float t12 = time_ms([&] { k_fpmul<BLS381><<<blocks, threads>>>(out, 3, iters); });
        printf("{\"bench\": \"fp_mul\", \"warps_per_sm\": %d, \"bn254_Gmul_per_s\": %.2f, \"bn254_wideMAC_T_per_s\": %.3f, \"bls381_Gmul_per_s\": %.2f, \"bls381_wideMAC_T_per_s\": %.3f}\n",
               warps * 2, muls / t8 / 1e6, muls * 136 / t8 / 1e9, muls / t12 / 1e6, muls * 300 / t12 / 1e9);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { printf("cuda error %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
