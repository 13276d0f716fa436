// One shape of tools/agg_bench.cu for ncu captures: 65,536 bls12-381 G2 points, the library's block shape.
//   ncu --set full --clock-control none --import-source on -k regex:k_agg -s 2 -c 1 -o gpurun_out/agg tools/agg_bench_one
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "../bgls_b200/csrc/agg.cuh"
using namespace bgls;
int main() {
    using E = AggFp2<BLS381>;
    constexpr int WPB = 4, NGB = WPB * AGG_GPW;
    const size_t n = 65536, rec = 192;
    std::vector<uint8_t> h(n * rec);
    srand(1);
    for (size_t i = 0; i < n * 4; i++) { uint8_t* f = h.data() + i * 48; for (int b = 0; b < 48; b++) f[b] = (uint8_t)rand(); f[0] &= 0x0f; }
    uint8_t *d_pts, *d_out; uint32_t* d_lv; unsigned* d_t;
    cudaMalloc(&d_pts, n * rec); cudaMemcpy(d_pts, h.data(), n * rec, cudaMemcpyHostToDevice); cudaMalloc(&d_out, rec);
    const size_t smem = agg_smem_bytes<E, NGB>();
    cudaFuncSetAttribute(k_agg<E, WPB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int occ = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_agg<E, WPB>, WPB * 32, smem);
    const size_t nb = 148 * (size_t)occ;
    cudaMalloc(&d_lv, (agg_tree_values(nb, NGB) + 1) * 72 * 4); cudaMalloc(&d_t, 4096 * 4); cudaMemset(d_t, 0, 4096 * 4);
    for (int i = 0; i < 4; i++) k_agg<E, WPB><<<(unsigned)nb, WPB * 32, smem>>>(d_pts, n, d_lv, d_t, d_out, nullptr);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
