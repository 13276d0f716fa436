// Microbenchmark of the six-lane aggregation kernel (bgls_b200/csrc/agg.cuh) outside the library: time per launch
// against points per launch and resident blocks per SM.  Inputs are random field elements below p (the complete
// formulas are data independent; correctness is covered by tests/test_agg_emul.py and the GPU parity tests).
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -o tools/agg_bench tools/agg_bench.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "../bgls_b200/csrc/agg.cuh"
using namespace bgls;

template <class E, int WPB, int MINB = 1> static void run(const char* name, size_t n, int per_sm_cap, int reps) {
    using C = typename E::Curve;
    constexpr int NGB = WPB * AGG_GPW, FB = C::FP_BYTES, NC = 2 * E::HALVES;
    const size_t rec = (size_t)NC * FB;
    std::vector<uint8_t> h(n * rec);
    srand(1);
    for (size_t i = 0; i < n * NC; i++) {
        uint8_t* f = h.data() + i * FB;
        for (int b = 0; b < FB; b++) f[b] = (uint8_t)rand();
        f[0] &= C::IS_BN ? 0x1f : 0x0f;   // below p
    }
    uint8_t *d_pts, *d_out;
    uint32_t* d_lv;
    unsigned* d_t;
    cudaMalloc(&d_pts, n * rec);
    cudaMemcpy(d_pts, h.data(), n * rec, cudaMemcpyHostToDevice);
    cudaMalloc(&d_out, rec);
    const size_t smem = agg_smem_bytes<E, NGB>();
    cudaFuncSetAttribute(k_agg<E, WPB, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int occ = 0, sms = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_agg<E, WPB, MINB>, WPB * 32, smem);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int per_sm = per_sm_cap > 0 && per_sm_cap < occ ? per_sm_cap : occ;
    size_t nb = (n + NGB - 1) / NGB;
    if (nb > (size_t)sms * per_sm) nb = (size_t)sms * per_sm;
    cudaMalloc(&d_lv, (agg_tree_values(nb, NGB) + 1) * 3 * E::HALVES * C::N * 4);
    unsigned long long* d_tr;
    cudaMalloc(&d_tr, 64);
    cudaMalloc(&d_t, 4096 * 4);
    cudaMemset(d_t, 0, 4096 * 4);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int i = 0; i < 3; i++) k_agg<E, WPB, MINB><<<(unsigned)nb, WPB * 32, smem>>>(d_pts, n, d_lv, d_t, d_out, d_tr);
    cudaDeviceSynchronize();
    float best = 1e9f, sum = 0;
    for (int i = 0; i < reps; i++) {
        cudaEventRecord(e0);
        k_agg<E, WPB, MINB><<<(unsigned)nb, WPB * 32, smem>>>(d_pts, n, d_lv, d_t, d_out, d_tr);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        best = ms < best ? ms : best;
        sum += ms;
    }
    cudaError_t err = cudaGetLastError();
    unsigned long long tr[5];
    cudaMemcpy(tr, d_tr, 40, cudaMemcpyDeviceToHost);
    printf("{\"kernel\": \"%s\", \"wpb\": %d, \"points\": %zu, \"blocks\": %zu, \"blocks_per_sm\": %d, \"occupancy_limit\": %d, \"smem\": %zu, "
           "\"ms_mean\": %.4f, \"ms_best\": %.4f, \"last_block_us\": {\"phase1\": %.1f, \"block_tree\": %.1f, \"cross_tree\": %.1f, \"affine\": %.1f}, \"minb\": %d, \"err\": \"%s\"}\n",
           name, WPB, n, nb, per_sm, occ, smem, sum / reps, best, (tr[1] - tr[0]) * 1e-3, (tr[2] - tr[1]) * 1e-3, (tr[3] - tr[2]) * 1e-3,
           (tr[4] - tr[3]) * 1e-3, MINB, cudaGetErrorString(err));
    cudaFree(d_pts); cudaFree(d_out); cudaFree(d_lv); cudaFree(d_t);
}

int main(int argc, char** argv) {
    const int reps = 20;
    for (int cap : {0, 1, 2}) {
        run<AggFp2<BLS381>, 4>("bls12-381 G2", 65536, cap, reps);
        run<AggFp2<BLS381>, 8>("bls12-381 G2", 65536, cap, reps);
    }
    run<AggFp2<BLS381>, 4, 4>("bls12-381 G2", 65536, 0, reps);
    run<AggFp2<BLS381>, 4, 4>("bls12-381 G2", 65536, 2, reps);
    run<AggFp2<BLS381>, 8, 2>("bls12-381 G2", 65536, 0, reps);
    run<AggFp2<BLS381>, 16, 1>("bls12-381 G2", 65536, 0, reps);
    for (size_t n : {2, 64, 1024, 8192, 262144}) run<AggFp2<BLS381>, 4>("bls12-381 G2", n, 0, reps);
    run<AggFp2<BLS381>, 8>("bls12-381 G2", 262144, 0, reps);
    run<AggFp2<BLS381>, 4, 4>("bls12-381 G2", 262144, 0, reps);
    run<AggFp<BLS381>, 4>("bls12-381 G1", 65536, 0, reps);
    run<AggFp2<BN254>, 4>("altbn128 G2", 65536, 0, reps);
    run<AggFp<BN254>, 4>("altbn128 G1", 65536, 0, reps);
    return 0;
}
