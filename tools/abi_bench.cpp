// Throughput of the C ABI without any Python in the loop: T host threads, each calling bgls_pairing_product (host buffers)
// or bgls_pairing_product_dev (device buffers, own stream) R times on n synthetic pairs (field elements below p; the
// verdict is irrelevant here, tests/ check it).      abi_bench <curve> <pairs> <threads> <reps> <dev 0|1>
#include <cuda_runtime.h>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
#include "../include/bgls_b200.h"

int main(int argc, char** argv) {
    const int curve = argc > 1 ? atoi(argv[1]) : 0;
    const size_t n = argc > 2 ? atoll(argv[2]) : 1025;
    const int T = argc > 3 ? atoi(argv[3]) : 16, R = argc > 4 ? atoi(argv[4]) : 50, dev = argc > 5 ? atoi(argv[5]) : 0;
    const size_t F = curve == 0 ? 32 : 48;
    bgls_ctx* ctx = nullptr;
    if (bgls_ctx_create(0, &ctx) != 0) { printf("ctx failed\n"); return 1; }
    std::vector<uint8_t> h1(n * 2 * F), h2(n * 4 * F);
    srand(3);
    for (auto& b : h1) b = rand() & 0xff;
    for (auto& b : h2) b = rand() & 0xff;
    for (size_t i = 0; i < n * 2; i++) h1[i * F] &= 0x0f;
    for (size_t i = 0; i < n * 4; i++) h2[i * F] &= 0x0f;
    uint8_t *p1, *p2;
    cudaHostAlloc((void**)&p1, h1.size(), 0); cudaHostAlloc((void**)&p2, h2.size(), 0);
    memcpy(p1, h1.data(), h1.size()); memcpy(p2, h2.data(), h2.size());
    uint8_t *d1, *d2;
    cudaMalloc((void**)&d1, h1.size()); cudaMalloc((void**)&d2, h2.size());
    cudaMemcpy(d1, p1, h1.size(), cudaMemcpyHostToDevice); cudaMemcpy(d2, p2, h2.size(), cudaMemcpyHostToDevice);
    auto body = [&](int t, int reps) {
        std::vector<uint8_t> gt(12 * F);
        int flag = 0;
        cudaStream_t st;
        cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
        uint8_t* dout; int* dflag;
        cudaMalloc((void**)&dout, 12 * F); cudaMalloc((void**)&dflag, 4);
        for (int r = 0; r < reps; r++) {
            if (dev == 3) {   // Miller product only (no final exponentiation), sync per call
                bgls_miller_product_dev(ctx, curve, d1, d2, n, dout, st);
                cudaStreamSynchronize(st);
            } else if (dev) {
                bgls_pairing_product_dev(ctx, curve, d1, d2, n, dout, dflag, st);
                if (dev == 2) cudaStreamSynchronize(st);
            } else {
                bgls_pairing_product(ctx, curve, p1, p2, n, gt.data(), &flag);
            }
        }
        cudaStreamSynchronize(st);
    };
    for (int phase = 0; phase < 2; phase++) {   // phase 0: warm-up
        const int reps = phase ? R : 3;
        auto t0 = std::chrono::steady_clock::now();
        std::vector<std::thread> th;
        for (int t = 0; t < T; t++) th.emplace_back(body, t, reps);
        for (auto& x : th) x.join();
        cudaDeviceSynchronize();
        const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        if (phase) printf("{\"curve\": %d, \"pairs\": %zu, \"threads\": %d, \"reps\": %d, \"mode\": \"%s\", \"ms_per_product\": %.4f, \"Mpairings_per_s\": %.3f}\n",
                          curve, n, T, R, dev == 0 ? "host buffers" : dev == 1 ? "device, enqueue only" : dev == 2 ? "device, sync per call" : "device, Miller product only, sync per call", dt / (T * R) * 1e3, (double)n * T * R / dt / 1e6);
    }
    bgls_ctx_destroy(ctx);
    return 0;
}
