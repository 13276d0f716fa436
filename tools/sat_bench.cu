// Microbenchmark of the slot engine's arithmetic core (bgls_b200/csrc/sat.cuh): one Fp2 operation per thread and
// iteration with operands in shared memory (the slot-file layout of slotvm.cuh), at 1 / 2 / 4 warps per SM
// sub-partition.  Reports 32x32->64 multiply-accumulates per second against the IMAD.WIDE roofline.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include "../bgls_b200/csrc/sat.cuh"
#include "../bgls_b200/csrc/field.cuh"
using namespace bgls;

template <class C, int MODE>
__global__ void __launch_bounds__(32) k(uint32_t* out, int iters) {
    constexpr int N = C::N, W4 = 2 * N / 4, NP = 32, NSLOT = 4;
    __shared__ uint4 sm[NSLOT * W4 * NP];
    const int q = threadIdx.x;
    for (int s = 0; s < NSLOT; s++)
        for (int w = 0; w < W4; w++) sm[(s * W4 + w) * NP + q] = make_uint4(q * 7 + s + 1, w * 13 + 5, s * 11 + q, (w + q) & 0xfffffff);
    __syncwarp();
    for (int it = 0; it < iters; it++) {
        const int sa = it & 1, sb = 2 + ((it >> 1) & 1), sd = (it + 1) & 1;
        uint32_t a[2 * N], b[2 * N], r[2 * N];
#pragma unroll
        for (int w = 0; w < W4; w++) {
            const uint4 x = sm[(sa * W4 + w) * NP + q], y = sm[(sb * W4 + w) * NP + q];
            a[4 * w] = x.x; a[4 * w + 1] = x.y; a[4 * w + 2] = x.z; a[4 * w + 3] = x.w & 0x0fffffffu;
            b[4 * w] = y.x; b[4 * w + 1] = y.y; b[4 * w + 2] = y.z; b[4 * w + 3] = y.w & 0x0fffffffu;
        }
        if (MODE == 0) sat_fp2_mul<C>(r, r + N, a, a + N, b, b + N);
        if (MODE == 1) sat_fp2_sqr<C>(r, r + N, a, a + N);
        if (MODE == 2) { mp_add<C>(r, a, b); mp_sub<C>(r + N, a + N, b + N); }
        if (MODE == 3) {   // round-1 arithmetic: three complete Montgomery multiplications (CIOS), canonical adds
            Fp2<C> fa, fb, fr;
#pragma unroll
            for (int i = 0; i < N; i++) { fa.c0.v[i] = a[i]; fa.c1.v[i] = a[N + i]; fb.c0.v[i] = b[i]; fb.c1.v[i] = b[N + i]; }
            fp2_mul(fr, fa, fb);
#pragma unroll
            for (int i = 0; i < N; i++) { r[i] = fr.c0.v[i]; r[N + i] = fr.c1.v[i]; }
        }
#pragma unroll
        for (int w = 0; w < W4; w++) sm[(sd * W4 + w) * NP + q] = make_uint4(r[4 * w], r[4 * w + 1], r[4 * w + 2], r[4 * w + 3]);
    }
    uint32_t acc = 0;
    for (int w = 0; w < W4; w++) acc ^= sm[w * NP + q].x;
    out[blockIdx.x * 32 + q] = acc;
}
template <class K> float time_ms(K launch) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    launch(); cudaDeviceSynchronize();
    cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
template <class C> void run(const char* name, int sms, uint32_t* out) {
    constexpr int N = C::N;
    const int iters = 2000;
    for (int wps = 1; wps <= 8; wps *= 2) {   // warps per SM sub-partition
        const int blocks = sms * 4 * wps;
        const double n = (double)blocks * 32 * iters;
        const float t0 = time_ms([&] { k<C, 0><<<blocks, 32>>>(out, iters); });
        const float t1 = time_ms([&] { k<C, 1><<<blocks, 32>>>(out, iters); });
        const float t2 = time_ms([&] { k<C, 2><<<blocks, 32>>>(out, iters); });
        const float t3 = time_ms([&] { k<C, 3><<<blocks, 32>>>(out, iters); });
        const double mac_mul = 3.0 * N * N + 2.0 * (N * N + N), mac_sqr = 2.0 * N * N + 2.0 * (N * N + N);
        printf("{\"curve\": \"%s\", \"warps_per_smsp\": %d, \"fp2_mul_G_per_s\": %.2f, \"fp2_mul_TMAC_per_s\": %.3f, \"fp2_sqr_G_per_s\": %.2f, "
               "\"fp2_sqr_TMAC_per_s\": %.3f, \"fp2_addsub_G_per_s\": %.2f, \"cycles_per_fp2_mul_per_warp\": %.0f, \"r1_cios_fp2_mul_G_per_s\": %.2f}\n",
               name, wps, n / t0 / 1e6, n * mac_mul / t0 / 1e9, n / t1 / 1e6, n * mac_sqr / t1 / 1e9, n / t2 / 1e6,
               t0 * 1e-3 * 1.965e9 / iters, n / t3 / 1e6);
    }
}
int main() {
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    uint32_t* out; cudaMalloc(&out, (size_t)prop.multiProcessorCount * 64 * 32 * 4);
    run<BN254>("altbn128", prop.multiProcessorCount, out);
    run<BLS381>("bls12-381", prop.multiProcessorCount, out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
