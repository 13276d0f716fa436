// Timing harness for the slot engine's Miller kernel (bgls_b200/csrc/slotvm.cuh): synthetic field elements
// (arithmetic is data independent), n pairs per product, S products in flight on S streams.
//   slot_bench <curve 0|1> <shape: 1|2|4 lanes per pair, 82 = 8 lanes per 2 pairs, 42 = 4 lanes per 2 pairs> <WPB 1|2|4> <pairs per product> <products in flight> [reps]
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../bgls_b200/csrc/slotvm.cuh"
using namespace bgls;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

template <class C, class T, int WPB>
int run(size_t n, int S, int reps) {
    constexpr int NPB = WPB * 32 / T::G;
    const size_t smem = sv_smem_bytes<C, T, NPB>() + (getenv("SLOT_EXTRA_SMEM") ? atoi(getenv("SLOT_EXTRA_SMEM")) : 0);
    CK(cudaFuncSetAttribute(k_slot_miller<C, T, WPB, SvNoFinish>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_slot_miller<C, T, WPB, SvNoFinish>, WPB * 32, smem));
    SvTables tb;
    uint32_t *d_code, *d_offs, *d_consts; uint8_t* d_seq;
    CK(cudaMalloc(&d_code, T::NWORDS * 4)); CK(cudaMemcpy(d_code, T::code(), T::NWORDS * 4, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&d_offs, (T::NPROG + 1) * 4)); CK(cudaMemcpy(d_offs, T::offsets(), (T::NPROG + 1) * 4, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&d_seq, T::SEQ_LEN)); CK(cudaMemcpy(d_seq, T::sequence(), T::SEQ_LEN, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&d_consts, T::NCONST * 2 * C::N * 4)); CK(cudaMemcpy(d_consts, T::consts(), T::NCONST * 2 * C::N * 4, cudaMemcpyHostToDevice));
    tb.code = d_code; tb.offs = d_offs; tb.seq = d_seq; tb.consts = d_consts; tb.mach_r = d_consts;
    const size_t FB = C::FP_BYTES;
    std::vector<uint8_t> h1(n * 2 * FB), h2(n * 4 * FB);
    srand(7);
    for (auto& b : h1) b = rand() & 0xff;
    for (auto& b : h2) b = rand() & 0xff;
    for (size_t i = 0; i < n * 2; i++) h1[i * FB] &= 0x0f;
    for (size_t i = 0; i < n * 4; i++) h2[i * FB] &= 0x0f;
    uint8_t *d1, *d2;
    CK(cudaMalloc(&d1, h1.size())); CK(cudaMalloc(&d2, h2.size()));
    CK(cudaMemcpy(d1, h1.data(), h1.size(), cudaMemcpyHostToDevice)); CK(cudaMemcpy(d2, h2.data(), h2.size(), cudaMemcpyHostToDevice));
    const unsigned nb = (unsigned)((n + (size_t)NPB * T::K - 1) / ((size_t)NPB * T::K));
    const size_t fan = 2 * NPB, lvw = sv_tree_words(nb, fan, C::N) + 16, ncnt = sv_tree_counters(nb, fan) + 16;
    uint32_t *dlv, *dmv; unsigned* dcnt;
    CK(cudaMalloc(&dlv, (size_t)S * lvw * 4)); CK(cudaMalloc(&dmv, (size_t)S * 12 * 16 * 4)); CK(cudaMalloc(&dcnt, (size_t)S * ncnt * 4));
    CK(cudaMemset(dcnt, 0, (size_t)S * ncnt * 4));
    std::vector<cudaStream_t> st(S);
    for (auto& s : st) CK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < reps + 1; rep++) {
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0, 0));
        for (int s = 0; s < S; s++) {
            CK(cudaStreamWaitEvent(st[s], e0, 0));
            k_slot_miller<C, T, WPB, SvNoFinish><<<nb, WPB * 32, smem, st[s]>>>(tb, d1, d2, n, dlv + (size_t)s * lvw, dcnt + (size_t)s * ncnt, dmv + (size_t)s * 12 * 16, 0, SvNoFinish::Args{0}, SvBatch{}, nullptr, SvPeers{});
        }
        CK(cudaGetLastError());
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e1, 0)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) best = ms;
    }
    printf("{\"curve\": \"%s\", \"G\": %d, \"K\": %d, \"WPB\": %d, \"pairs\": %zu, \"in_flight\": %d, \"blocks_per_product\": %u, \"smem_per_block\": %zu, "
           "\"blocks_per_sm\": %d, \"ms\": %.4f, \"Mpairings_per_s\": %.3f}\n",
           C::IS_BN ? "altbn128" : "bls12-381", T::G, T::K, WPB, n, S, nb, smem, occ, best, (double)n * S / best / 1e3);
    return 0;
}
template <class C, class T> int run_w(int wpb, size_t n, int S, int reps) {
    if (wpb == 1) return run<C, T, 1>(n, S, reps);
    if (wpb == 2) return run<C, T, 2>(n, S, reps);
    return run<C, T, 4>(n, S, reps);
}
int main(int argc, char** argv) {
    if (argc < 6) { printf("usage: slot_bench curve G WPB pairs in_flight [reps]\n"); return 2; }
    const int curve = atoi(argv[1]), g = atoi(argv[2]), wpb = atoi(argv[3]);
    const size_t n = (size_t)atoll(argv[4]);
    const int S = atoi(argv[5]), reps = argc > 6 ? atoi(argv[6]) : 3;
    if (curve == 0) {
        if (g == 1) return run_w<BN254, svt::BN254_G1>(wpb, n, S, reps);
        if (g == 2) return run_w<BN254, svt::BN254_G2>(wpb, n, S, reps);
        if (g == 82) return run_w<BN254, svt::BN254_G8K2>(wpb, n, S, reps);
        if (g == 8) return run_w<BN254, svt::BN254_G8>(wpb, n, S, reps);
        if (g == 16) return run_w<BN254, svt::BN254_G16>(wpb, n, S, reps);
        if (g == 42) return run_w<BN254, svt::BN254_G4K2>(wpb, n, S, reps);
        return run_w<BN254, svt::BN254_G4>(wpb, n, S, reps);
    }
    if (g == 1) return run_w<BLS381, svt::BLS381_G1>(wpb, n, S, reps);
    if (g == 2) return run_w<BLS381, svt::BLS381_G2>(wpb, n, S, reps);
    if (g == 82) return run_w<BLS381, svt::BLS381_G8K2>(wpb, n, S, reps);
    if (g == 8) return run_w<BLS381, svt::BLS381_G8>(wpb, n, S, reps);
    if (g == 16) return run_w<BLS381, svt::BLS381_G16>(wpb, n, S, reps);
    if (g == 42) return run_w<BLS381, svt::BLS381_G4K2>(wpb, n, S, reps);
    return run_w<BLS381, svt::BLS381_G4>(wpb, n, S, reps);
}
