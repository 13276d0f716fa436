// bgls_b200: CUDA kernels (sm_100a) and the C ABI declared in include/bgls_b200.h.
//
// Kernel inventory:
//   k_miller_product   K1+K2  one Miller loop per thread, per-block Fp12 product tree in SMEM
//   k_finish           K2+K3  product of the per-block partials, one final exponentiation
//   k_miller_values / k_batch_finish   throughput mode: many independent products per launch
//   k_aggregate / k_aggregate_finish   K4  n-way G1/G2 sum (AggregatePoints)
//   k_scale            ScalePoints (Point.Mul)
// There is no CPU fallback: without a CUDA device bgls_ctx_create fails with BGLS_ERR_NODEV.
#include <cuda_runtime.h>

#include <cstdio>
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <chrono>
#include <mutex>
#include <string>
#include <unordered_set>
#include <vector>

#include "../../include/bgls_b200.h"
#include "pairing.cuh"
#include "machine_kernels.cuh"
#include "slotvm.cuh"
#include "agg.cuh"
#include "hash.cuh"
#include "codec.cuh"

namespace bgls {

constexpr int TB = 32;        // threads per block, thread-per-item kernels
constexpr int MAX_BLOCKS = 148 * 16;

// ---------------------------------------------------------------- block-wide Fp12 product
template <class C> __device__ void block_product(Fp12<C>& acc, Fp12<C>* sh) {
    const int tid = threadIdx.x;
    sh[tid] = acc;
    __syncthreads();
    for (int s = TB / 2; s > 0; s >>= 1) {
        if (tid < s) {
            fp12_mul(acc, acc, sh[tid + s]);
            sh[tid] = acc;
        }
        __syncthreads();
    }
}

template <class C>
__global__ void __launch_bounds__(TB) k_miller_product(const uint8_t* __restrict__ g1, const uint8_t* __restrict__ g2,
                                                       size_t n, Fp12<C>* __restrict__ partial) {
    __shared__ Fp12<C> sh[TB];
    Fp12<C> acc;
    fp12_one(acc);
    bool first = true;
    for (size_t i = (size_t)blockIdx.x * TB + threadIdx.x; i < n; i += (size_t)gridDim.x * TB) {
        G1Aff<C> P;
        G2Aff<C> Q;
        g1_load<C>(P, g1 + i * 2 * C::FP_BYTES);
        g2_load<C>(Q, g2 + i * 4 * C::FP_BYTES);
        if (first) {
            miller_loop(acc, P, Q);
            first = false;
        } else {
            Fp12<C> f;
            miller_loop(f, P, Q);
            fp12_mul(acc, acc, f);
        }
    }
    block_product(acc, sh);
    if (threadIdx.x == 0) partial[blockIdx.x] = acc;
}

// Throughput variant: every thread carries K consecutive pairs through one shared Miller accumulator
// (miller_loop_shared): one Fp12 squaring per iteration for K pairs.
constexpr int MSHARE = 4;
constexpr size_t MSHARED_MIN = 32768;   // pairs: below this one pair per thread keeps more warps busy
template <class C, int K>
__global__ void __launch_bounds__(TB) k_miller_product_shared(const uint8_t* __restrict__ g1, const uint8_t* __restrict__ g2,
                                                              size_t n, Fp12<C>* __restrict__ partial) {
    __shared__ Fp12<C> sh[TB];
    Fp12<C> acc;
    fp12_one(acc);
    bool first = true;
    const size_t groups = (n + K - 1) / K;
    for (size_t g = (size_t)blockIdx.x * TB + threadIdx.x; g < groups; g += (size_t)gridDim.x * TB) {
        G1Aff<C> P[K];
        G2Aff<C> Q[K];
        const size_t base = g * K;
        const int k = (int)(n - base < (size_t)K ? n - base : (size_t)K);
        for (int j = 0; j < k; j++) {
            g1_load<C>(P[j], g1 + (base + j) * 2 * C::FP_BYTES);
            g2_load<C>(Q[j], g2 + (base + j) * 4 * C::FP_BYTES);
        }
        if (first) {
            miller_loop_shared<C, K>(acc, P, Q, k);
            first = false;
        } else {
            Fp12<C> f;
            miller_loop_shared<C, K>(f, P, Q, k);
            fp12_mul(acc, acc, f);
        }
    }
    block_product(acc, sh);
    if (threadIdx.x == 0) partial[blockIdx.x] = acc;
}

// Batch of independent products (throughput mode): one block per check, every thread carries up to KB_MAX pairs of
// the check through one shared Miller accumulator; the block product leaves the check's raw Miller product as wire bytes.
constexpr int TBB = 64;
constexpr int KB_MAX = 8;
template <class C>
__global__ void __launch_bounds__(TBB) k_miller_batch_shared(const uint8_t* __restrict__ g1, const uint8_t* __restrict__ g2,
                                                            const unsigned long long* __restrict__ off, size_t nbatch,
                                                            uint8_t* __restrict__ raw) {
    __shared__ Fp12<C> sh[TBB];
    const size_t b = blockIdx.x;
    const size_t lo = off[b], hi = off[b + 1];
    size_t k = (hi - lo + TBB - 1) / TBB;
    if (k > (size_t)KB_MAX) k = KB_MAX;
    if (k == 0) k = 1;
    Fp12<C> acc;
    fp12_one(acc);
    bool first = true;
    for (size_t base = lo + (size_t)threadIdx.x * k; base < hi; base += (size_t)TBB * k) {
        G1Aff<C> P[KB_MAX];
        G2Aff<C> Q[KB_MAX];
        const int kk = (int)(hi - base < k ? hi - base : k);
        for (int j = 0; j < kk; j++) {
            g1_load<C>(P[j], g1 + (base + j) * 2 * C::FP_BYTES);
            g2_load<C>(Q[j], g2 + (base + j) * 4 * C::FP_BYTES);
        }
        if (first) {
            miller_loop_shared<C, KB_MAX>(acc, P, Q, kk);
            first = false;
        } else {
            Fp12<C> f;
            miller_loop_shared<C, KB_MAX>(f, P, Q, kk);
            fp12_mul(acc, acc, f);
        }
    }
    const int tid = threadIdx.x;
    sh[tid] = acc;
    __syncthreads();
    for (int s = TBB / 2; s > 0; s >>= 1) {
        if (tid < s) {
            fp12_mul(acc, acc, sh[tid + s]);
            sh[tid] = acc;
        }
        __syncthreads();
    }
    if (tid == 0) fp12_to_be<C>(raw + b * 12 * C::FP_BYTES, acc);
}

// one block: multiply k partial values, optionally exponentiate, emit wire bytes + identity flag
template <class C, bool IN_BYTES>
__global__ void __launch_bounds__(TB) k_finish(const void* __restrict__ in, size_t k, int do_final,
                                               uint8_t* __restrict__ out_gt, int* __restrict__ is_identity) {
    __shared__ Fp12<C> sh[TB];
    Fp12<C> acc;
    fp12_one(acc);
    for (size_t i = threadIdx.x; i < k; i += TB) {
        Fp12<C> f;
        if (IN_BYTES) fp12_from_be<C>(f, (const uint8_t*)in + i * 12 * C::FP_BYTES);
        else f = ((const Fp12<C>*)in)[i];
        fp12_mul(acc, acc, f);
    }
    block_product(acc, sh);
    if (threadIdx.x == 0) {
        if (do_final) final_exp(acc, acc);
        fp12_to_be<C>(out_gt, acc);
        if (is_identity) *is_identity = fp12_is_one(acc) ? 1 : 0;
    }
}

template <class C>
__global__ void __launch_bounds__(TB) k_miller_values(const uint8_t* __restrict__ g1, const uint8_t* __restrict__ g2,
                                                      size_t n, Fp12<C>* __restrict__ vals) {
    const size_t i = (size_t)blockIdx.x * TB + threadIdx.x;
    if (i >= n) return;
    G1Aff<C> P;
    G2Aff<C> Q;
    Fp12<C> f;
    g1_load<C>(P, g1 + i * 2 * C::FP_BYTES);
    g2_load<C>(Q, g2 + i * 4 * C::FP_BYTES);
    miller_loop(f, P, Q);
    vals[i] = f;
}
template <class C>
__global__ void __launch_bounds__(TB) k_batch_finish(const Fp12<C>* __restrict__ vals, const uint64_t* __restrict__ offsets,
                                                     size_t nbatch, uint8_t* __restrict__ out_ok) {
    const size_t b = (size_t)blockIdx.x * TB + threadIdx.x;
    if (b >= nbatch) return;
    Fp12<C> acc;
    fp12_one(acc);
    for (uint64_t i = offsets[b]; i < offsets[b + 1]; i++) fp12_mul(acc, acc, vals[i]);
    final_exp(acc, acc);
    out_ok[b] = fp12_is_one(acc) ? 1 : 0;
}

// ---------------------------------------------------------------- point aggregation / scaling
template <class C, class F>
__global__ void __launch_bounds__(TB) k_aggregate(const uint8_t* __restrict__ pts, size_t n, size_t rec,
                                                  Jac<F>* __restrict__ partial) {
    __shared__ Jac<F> sh[TB];
    Jac<F> acc;
    acc.inf = true;
    for (size_t i = (size_t)blockIdx.x * TB + threadIdx.x; i < n; i += (size_t)gridDim.x * TB) {
        Jac<F> p;
        jac_load<C>(p, pts + i * rec);
        jac_add(acc, acc, p);
    }
    const int tid = threadIdx.x;
    sh[tid] = acc;
    __syncthreads();
    for (int s = TB / 2; s > 0; s >>= 1) {
        if (tid < s) {
            jac_add(acc, acc, sh[tid + s]);
            sh[tid] = acc;
        }
        __syncthreads();
    }
    if (tid == 0) partial[blockIdx.x] = acc;
}
// one level of the partial-sum tree: TB Jacobian partials per block -> one
template <class F>
__global__ void __launch_bounds__(TB) k_aggregate_level(const Jac<F>* __restrict__ in, size_t k, Jac<F>* __restrict__ out) {
    __shared__ Jac<F> sh[TB];
    const int tid = threadIdx.x;
    const size_t i = (size_t)blockIdx.x * TB + tid;
    Jac<F> acc;
    if (i < k) acc = in[i];
    else acc.inf = true;
    sh[tid] = acc;
    __syncthreads();
    for (int s = TB / 2; s > 0; s >>= 1) {
        if (tid < s) {
            jac_add(acc, acc, sh[tid + s]);
            sh[tid] = acc;
        }
        __syncthreads();
    }
    if (tid == 0) out[blockIdx.x] = acc;
}
template <class C, class F>
__global__ void __launch_bounds__(TB) k_aggregate_finish(const Jac<F>* __restrict__ partial, size_t k, uint8_t* __restrict__ out) {
    __shared__ Jac<F> sh[TB];
    Jac<F> acc;
    acc.inf = true;
    for (size_t i = threadIdx.x; i < k; i += TB) jac_add(acc, acc, partial[i]);
    const int tid = threadIdx.x;
    sh[tid] = acc;
    __syncthreads();
    for (int s = TB / 2; s > 0; s >>= 1) {
        if (tid < s) {
            jac_add(acc, acc, sh[tid + s]);
            sh[tid] = acc;
        }
        __syncthreads();
    }
    if (tid == 0) jac_store<C>(out, acc);
}
template <class C, class F>
__global__ void __launch_bounds__(TB) k_scale(const uint8_t* __restrict__ pts, const uint8_t* __restrict__ scalars, size_t n,
                                              size_t rec, uint8_t* __restrict__ out) {
    const size_t i = (size_t)blockIdx.x * TB + threadIdx.x;
    if (i >= n) return;
    Jac<F> p, r;
    jac_load<C>(p, pts + i * rec);
    jac_mul(r, p, scalars + 32 * i);
    jac_store<C>(out + i * rec, r);
}

// hash-to-G1 (messages concatenated, offsets[n+1]); both kernels spread one message over several lanes because the
// per-message work is a chain of ~570-multiplication exponentiations:
//   altbn128  -- 8 lanes per message try 8 consecutive counters of the try-and-increment loop at once, the lowest
//                successful counter wins (what the sequential loop of hash.go:53-77 would have found first)
//   bls12-381 -- 2 lanes per message, one per half (G1_0 / G1_1) of the Fouque-Tibouchi hash, summed by the even lane
template <class C>
__global__ void __launch_bounds__(TB) k_hash_to_g1_bn(const uint8_t* __restrict__ msgs, const unsigned long long* __restrict__ off, size_t n,
                                                      uint8_t* __restrict__ out) {
    constexpr int LPM = 8;
    const size_t i = ((size_t)blockIdx.x * TB + threadIdx.x) / LPM;
    const int sub = threadIdx.x % LPM, grp = (threadIdx.x & 31) / LPM;
    const bool active = i < n;
    const uint8_t* m = active ? msgs + off[i] : msgs;
    const size_t len = active ? (size_t)(off[i + 1] - off[i]) : 0;
    bool done = !active;
    for (int round = 0; round < 256 / LPM; round++) {
        Fp<C> px, root;
        const bool ok = !done && bn_hash_try<C>(px, root, (uint8_t)(round * LPM + sub), m, len);
        const unsigned hits = (__ballot_sync(0xFFFFFFFFu, ok) >> (grp * LPM)) & ((1u << LPM) - 1);
        if (!done && hits) {
            if (sub == __ffs(hits) - 1) bn_hash_finish<C>(out + i * 2 * C::FP_BYTES, px, root, m, len);
            done = true;
        }
        if (__all_sync(0xFFFFFFFFu, done)) break;
    }
}
// altbn128, throughput form: a warp owns 32 messages and deals its lanes out again every round -- lane L tests counter
// next[m] + L / u of the (L mod u)-th unfinished message m, u = messages still open.  Round 1 is one counter per message
// (half succeed), round 2 two counters for each of the ~16 left, round 3 eight for the ~4 left.  The test of a counter is
// a Jacobi symbol (inv.cuh), not the square root: the one exponentiation per message runs after the search, on the
// lane that owns the message -- ~3 cheap tests + 1 exponentiation per message instead of the 8 exponentiations of the
// latency form above.  The lowest successful counter of a message wins, exactly what the sequential loop of
// hash.go:53-77 finds first.
template <class C>
__global__ void __launch_bounds__(TB) k_hash_to_g1_bn_pool(const uint8_t* __restrict__ msgs, const unsigned long long* __restrict__ off, size_t n,
                                                           uint8_t* __restrict__ out) {
    static_assert(TB == 32, "one warp per block");
    const int lane = threadIdx.x;
    const size_t base = (size_t)blockIdx.x * 32;
    const bool mine = base + lane < n;
    bool done = !mine;
    unsigned next = 0, won = 0;                          // next counter to test / the counter found, of this lane's own message
    for (;;) {
        const unsigned pending = __ballot_sync(0xFFFFFFFFu, !done);
        if (!pending) break;
        const int u = __popc(pending), per = 32 / u;     // `per` counters per open message this round
        const int slot = lane % u, rep = lane / u;
        const int owner = __fns(pending, 0, slot + 1);   // lane that owns the slot-th open message
        const unsigned counter = __shfl_sync(0xFFFFFFFFu, next, owner) + rep;
        const size_t i = base + owner;
        const bool ok = rep < per && counter < 256 && bn_hash_test<C>((uint8_t)counter, msgs + off[i], (size_t)(off[i + 1] - off[i]));
        const unsigned hits = __ballot_sync(0xFFFFFFFFu, ok);
        if (!done) {
            const int myslot = __popc(pending & ((1u << lane) - 1u));
            for (int r = 0; r < per && !done; r++)
                if ((hits >> (myslot + r * u)) & 1u) { won = next + r; done = true; }
            if (!done && next + per >= 256) { won = 256; done = true; }   // 256 failures: probability 2^-256, the reference would spin
            next += per;
        }
    }
    if (mine && won < 256) {
        const size_t i = base + lane;
        const uint8_t* m = msgs + off[i];
        const size_t len = (size_t)(off[i + 1] - off[i]);
        Fp<C> px, root;
        bn_hash_try<C>(px, root, (uint8_t)won, m, len);
        bn_hash_finish<C>(out + i * 2 * C::FP_BYTES, px, root, m, len);
    }
}
template <class C>
__global__ void __launch_bounds__(TB) k_hash_to_g1_bls(const uint8_t* __restrict__ msgs, const unsigned long long* __restrict__ off, size_t n,
                                                       uint8_t* __restrict__ out) {
    __shared__ Jac<Fp<C>> sh[TB / 2];
    const size_t i = ((size_t)blockIdx.x * TB + threadIdx.x) / 2;
    const int half = threadIdx.x & 1;
    Jac<Fp<C>> P;
    P.inf = true;
    if (i < n) ft_half<C>(P, msgs + off[i], (size_t)(off[i + 1] - off[i]), half);
    if (half) sh[threadIdx.x / 2] = P;
    __syncwarp();
    if (!half && i < n) {
        Jac<Fp<C>> S;
        jac_add(S, P, sh[threadIdx.x / 2]);
        jac_store<C>(out + i * 2 * C::FP_BYTES, S);
    }
}

// bls12-381, throughput form: one thread per message, one cofactor multiplication per message (hash.cuh)
template <class C>
__global__ void __launch_bounds__(TB) k_hash_to_g1_bls_one(const uint8_t* __restrict__ msgs, const unsigned long long* __restrict__ off, size_t n,
                                                           uint8_t* __restrict__ out) {
    const size_t i = (size_t)blockIdx.x * TB + threadIdx.x;
    if (i < n) hash_to_g1_ft_shared_cofactor<C>(out + i * 2 * C::FP_BYTES, msgs + off[i], (size_t)(off[i + 1] - off[i]));
}

// ---------------------------------------------------------------- peer-memory exchange of the per-GPU partials
// Every rank stores its 12F-byte Miller product straight into a mailbox on each peer (NVLink peer memory mapped
// through CUDA IPC) and then raises that mailbox's epoch flag; the finishing side waits for the flags of all ranks.
// Slots are double-buffered by epoch parity (a rank can be one step ahead on the same lane, never two).
template <class PP>
__global__ void __launch_bounds__(256) k_exchange_send_t(const uint8_t* __restrict__ partial, int nbytes, PP pp, int world, unsigned long long epoch) {
    for (int idx = threadIdx.x; idx < world * nbytes; idx += blockDim.x) {
        const int r = idx / nbytes, b = idx % nbytes;
        pp.slot[r][b] = partial[b];
    }
    __threadfence_system();   // the payload is visible to every peer before any flag is
    __syncthreads();
    if ((int)threadIdx.x < world) {
        *(volatile unsigned long long*)pp.flag[threadIdx.x] = epoch;
        __threadfence_system();
    }
}
// one thread per rank polls its flag; bounded, so a missing peer reports an error instead of hanging the GPU
__global__ void __launch_bounds__(32) k_exchange_wait(const unsigned long long* flags, int world, unsigned long long epoch,
                                                     int* __restrict__ err) {
    if ((int)threadIdx.x < world) {
        const volatile unsigned long long* f = flags + threadIdx.x;
        long long spins = 0;
        while (*f < epoch) {
            if (++spins > (1ll << 27)) { *err = 1; break; }   // several seconds
            __nanosleep(64);
        }
    }
    __threadfence_system();
}

// a wait that timed out must not leave a verdict computed from stale mailbox records behind
__global__ void k_exchange_guard(const int* __restrict__ err, int* __restrict__ is_identity) {
    if (*err && is_identity) *is_identity = 0;
}

// compressed wire formats, one point per thread (codec.cuh)
template <class C, int G>
__global__ void __launch_bounds__(TB) k_compress(const uint8_t* __restrict__ pts, size_t n, uint8_t* __restrict__ out) {
    const size_t i = (size_t)blockIdx.x * TB + threadIdx.x;
    if (i >= n) return;
    constexpr int F = C::FP_BYTES;
    if (G == 1) compress_g1<C>(out + i * F, pts + i * 2 * F);
    else compress_g2<C>(out + i * 2 * F, pts + i * 4 * F);
}
template <class C, int G>
__global__ void __launch_bounds__(TB) k_decompress(const uint8_t* __restrict__ in, size_t n, int check_subgroup,
                                                   uint8_t* __restrict__ pts, uint8_t* __restrict__ ok) {
    const size_t i = (size_t)blockIdx.x * TB + threadIdx.x;
    if (i >= n) return;
    constexpr int F = C::FP_BYTES;
    bool good;
    if (G == 1) good = decompress_g1<C>(pts + i * 2 * F, in + i * F, check_subgroup != 0);
    else good = decompress_g2<C>(pts + i * 4 * F, in + i * 2 * F, check_subgroup != 0);
    ok[i] = good ? 1 : 0;
}

template <class C, int G>
__global__ void __launch_bounds__(TB) k_validate(const uint8_t* __restrict__ pts, size_t n, int subgroup, uint8_t* __restrict__ ok) {
    const size_t i = (size_t)blockIdx.x * TB + threadIdx.x;
    if (i >= n) return;
    constexpr int F = C::FP_BYTES;
    ok[i] = (G == 1 ? validate_g1<C>(pts + i * 2 * F, subgroup != 0) : validate_g2<C>(pts + i * 4 * F, subgroup != 0)) ? 1 : 0;
}
template <class C>
__global__ void __launch_bounds__(32) k_gt_pow(const uint8_t* __restrict__ a, const uint8_t* __restrict__ e32, int negative, uint8_t* __restrict__ out) {
    if (threadIdx.x == 0) gt_pow<C>(out, a, e32, negative != 0);
}

// register-only IMAD.WIDE.U32 loop: the integer-pipe roofline denominator.  The multiplier operand is
// data dependent (low word of another accumulator), otherwise ptxas hoists the loop-invariant
// product and the loop degenerates into 64-bit additions (which is what a first version measured).
__global__ void k_intpipe_peak(uint32_t* out, uint32_t seed, int iters) {
    uint32_t b = seed * 3 + 1 + threadIdx.x;
    unsigned long long w[8];
#pragma unroll
    for (int i = 0; i < 8; i++) w[i] = (unsigned long long)(seed + threadIdx.x) * 7 + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const uint32_t a = (uint32_t)w[(i + 4) & 7];  // depends on the accumulator updated four instructions ago
            asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(a), "r"(b));
        }
    }
    unsigned long long s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += w[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = (uint32_t)s ^ (uint32_t)(s >> 32);
}

}  // namespace bgls

// ==================================================================== host side / C ABI
using namespace bgls;

// device copies of the machine tables of one curve
struct MachDev {
    MachTables m{}, f{}, p{};   // 16-lane product file, final-exponentiation file, pipelined 32-lane Miller file
    const uint16_t *miller = nullptr, *mul_ab = nullptr, *mul_ba = nullptr, *import_a = nullptr, *miller_p = nullptr, *mulacc_p = nullptr;
    const uint16_t *finalexp = nullptr, *export_ = nullptr;
    void* blob = nullptr;
};

// One execution slot: a stream with its own device scratch.  Host-buffer calls take any free slot, so calls made
// from several host threads (the reference calls PairingProduct from many goroutines, curves/curve.go:132-134)
// overlap on the GPU: the single-warp final exponentiation of one product runs beside the Miller loops of the next.
// Device-resident calls are keyed by the caller's stream, so work enqueued on different streams never shares scratch.
struct Slot {
    cudaStream_t stream = nullptr;   // the slot's own stream (host-buffer calls)
    cudaStream_t aux = nullptr;      // second stream for independent work inside one call (fork / join events below)
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    cudaEvent_t ev_done = nullptr;   // recorded on the caller's stream at the end of every device-resident call using this slot
    std::atomic<long long> last_use_us{0};   // time of the last pairing call that used this slot (load estimate)
    cudaStream_t owner = nullptr;    // caller stream this slot's scratch is currently ordered on (device-resident calls)
    bool owned = false;
    void* scratch = nullptr;
    size_t scratch_bytes = 0;
    unsigned* tickets = nullptr;     // slot engine: tickets of the in-launch product tree (zero between launches)
    size_t tickets_cap = 0;
    uint8_t* hres_dev = nullptr;     // device address of `hres` (mapped): the kernels write the small results straight into it
    uint8_t* hres = nullptr;         // pinned host buffer for the small results (GT bytes + verdict): the device-to-host
                                     // copy of a call never goes through the driver's pageable staging path
    std::mutex mu;
};
constexpr size_t HRES_BYTES = 1024;
constexpr int NSLOT = 64;

// peer-memory exchange state (bgls_exchange_*): this rank's mailbox and the mapped mailboxes of the peers
struct Exchange {
    int world = 0, rank = 0, lanes = 0;
    uint8_t* local = nullptr;                 // [lanes][2][world] records of XREC bytes, then [lanes][2][world] flags
    std::vector<uint8_t*> peer;               // base pointers of every rank's mailbox (own rank: local)
    uint8_t** d_slot = nullptr;               // device tables [lanes][2][world]: where MY record lives in rank r's mailbox
    unsigned long long** d_flag = nullptr;
    int* d_err = nullptr;
    bool ready = false;
};
constexpr size_t XREC = 640;                  // 12 * 48 bytes rounded up to a multiple of 128

// device copies of the slot engine's tables of one curve (slotvm.cuh)
struct SlotEngDev {
    SvTables tb{};
    void* blob = nullptr;
};
enum { ENGINE_AUTO = 0, ENGINE_MACHINE, ENGINE_THREAD, ENGINE_SLOT };

struct bgls_ctx {
    MachDev mach[2];
    SlotEngDev sloteng[2];
    int engine = ENGINE_AUTO;        // BGLS_ENGINE=auto|machine|thread|slot
    unsigned long long* trace = nullptr;   // BGLS_TRACE=file: per-block timeline of the slot Miller kernel, dumped at destroy
    std::atomic<int> host_calls{0};  // host-buffer pairing calls currently inside the library
    Exchange xch;
    Slot slots[NSLOT];
    std::atomic<unsigned> rr{0};
    std::atomic<size_t> scratch_hint{0};
    std::mutex own_mu;
    bool thread_engine = false;  // BGLS_ENGINE=thread: thread-per-pair kernels only
    bool machine_only = false;   // BGLS_ENGINE=machine: dot-product machine at every size (no hybrid)
    bool no_shared = false;      // BGLS_MILLER=noshare: thread engine without the shared Miller accumulator (one pair per thread)
    int slot_split = -1;         // BGLS_SLOT_SPLIT=0|1: final exponentiation of the slot pipeline inside the Miller launch / as its own
                                 // launch; default (-1): own launch on altbn128, inside on bls12-381 (measured, profiles/r2_aa_*)
    int hash_mode = 0;           // BGLS_HASH=pool|wide: altbn128 try-and-increment form (0: by load)
    int agg_blocks_per_sm = 0;   // BGLS_AGG_BLOCKS=k: at most k aggregation blocks per SM (0: as many as fit)
    int min_wpb = 8;             // BGLS_MIN_WPB=k: at least k pairings (warps) per block of k_mach_miller32
    bool miller16 = false;       // BGLS_MILLER=m16: 16-lane Miller program (two pairings per warp) instead of the pipelined one
    int device = 0;
    int sms = 148;
    std::atomic<uint64_t> launches{0};
    bool profiling = false;      // kernel-time events: single-threaded use only
    cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};  // before main kernel, between, after finish
    std::mutex err_mu;
    std::string err;
};

namespace {

const char* kVersion = "bgls_b200 r1 sm_100a";

int fail(bgls_ctx* ctx, int code, const char* what, cudaError_t e = cudaSuccess) {
    if (ctx) {
        std::lock_guard<std::mutex> lk(ctx->err_mu);
        ctx->err = what;
        if (e != cudaSuccess) { ctx->err += ": "; ctx->err += cudaGetErrorString(e); }
    }
    return code;
}
#define CU(call)                                                         \
    do {                                                                 \
        cudaError_t e__ = (call);                                        \
        if (e__ != cudaSuccess) return fail(ctx, BGLS_ERR_CUDA, #call, e__); \
    } while (0)

size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

int ensure_scratch(bgls_ctx* ctx, Slot* sl, size_t bytes) {
    if (bytes <= sl->scratch_bytes) return BGLS_OK;
    // GROWTH SYNCHRONISES (documented in include/bgls_b200.h): the old buffer may still be in use by work enqueued earlier,
    // so the first call of a given size on a slot waits for the device; steady-state calls never get here.
    // Not legal inside a stream capture: warm the context up with one call of the largest size before capturing.
    CU(cudaDeviceSynchronize());
    if (sl->scratch) CU(cudaFree(sl->scratch));
    sl->scratch = nullptr;
    sl->scratch_bytes = 0;
    // all slots converge on the largest request seen by any of them: a slot grows at most once per new maximum
    size_t want = align_up(bytes + bytes / 4, 1 << 20);
    size_t hint = ctx->scratch_hint.load();
    while (hint < want && !ctx->scratch_hint.compare_exchange_weak(hint, want)) {}
    if (hint > want) want = hint;
    CU(cudaMalloc(&sl->scratch, want));
    sl->scratch_bytes = want;
    return BGLS_OK;
}
int blocks_for(size_t n) {
    size_t b = (n + TB - 1) / TB;
    if (b < 1) b = 1;
    if (b > (size_t)MAX_BLOCKS) b = MAX_BLOCKS;
    return (int)b;
}
size_t jac_dev_bytes(int c, int g);
// scratch of the aggregation pipeline: block partials + the second buffer of the level tree
size_t agg_work_bytes(int curve, int group, size_t n);
bool curve_ok(int c) { return c == BGLS_ALTBN128 || c == BGLS_BLS12_381; }
size_t fp_bytes(int c) { return c == BGLS_ALTBN128 ? 32 : 48; }
size_t fp12_dev_bytes(int c) { return c == BGLS_ALTBN128 ? sizeof(Fp12<BN254>) : sizeof(Fp12<BLS381>); }
size_t jac_dev_bytes(int c, int g) {
    if (c == BGLS_ALTBN128) return g == 1 ? sizeof(Jac<Fp<BN254>>) : sizeof(Jac<Fp2<BN254>>);
    return g == 1 ? sizeof(Jac<Fp<BLS381>>) : sizeof(Jac<Fp2<BLS381>>);
}

// ---- enqueue helpers (device pointers); `work` is device scratch owned by the caller
size_t agg_work_bytes(int curve, int group, size_t n) {
    // the grid never exceeds SMs x resident blocks per SM (agg_grid): 4096 bounds it on any part
    const size_t nb = std::min<size_t>((n + 19) / 20 + 1, 4096), words = 3 * (size_t)group * (curve == BGLS_ALTBN128 ? 8 : 12);
    return align_up((agg_tree_values(nb, 20) + 1) * words * 4);
}
template <class C>
int enqueue_pairing(bgls_ctx* ctx, const uint8_t* d_g1, const uint8_t* d_g2, size_t n, int do_final, uint8_t* d_out,
                    int* d_flag, void* work, cudaStream_t s, bool prof) {
    const int nb = blocks_for(n);
    int nb_used = nb;
    Fp12<C>* partial = (Fp12<C>*)work;
    if (prof) cudaEventRecord(ctx->ev[0], s);
    if (n >= MSHARED_MIN && !ctx->no_shared) {
        const int nbs = blocks_for((n + MSHARE - 1) / MSHARE);
        k_miller_product_shared<C, MSHARE><<<nbs, TB, 0, s>>>(d_g1, d_g2, n, partial);
        ctx->launches++;
        nb_used = nbs;
    } else if (n > 0) {
        k_miller_product<C><<<nb, TB, 0, s>>>(d_g1, d_g2, n, partial);
        ctx->launches++;
    }
    if (prof) cudaEventRecord(ctx->ev[1], s);
    k_finish<C, false><<<1, TB, 0, s>>>(partial, n > 0 ? (size_t)nb_used : 0, do_final, d_out, d_flag);
    ctx->launches++;
    if (prof) cudaEventRecord(ctx->ev[2], s);
    CU(cudaGetLastError());
    return BGLS_OK;
}
// AggregatePoints in one launch (agg.cuh): six lanes per addition, in-launch block and cross-block trees, one binary
// inversion.  `work` = levels of the cross-block tree (agg_work_bytes); the tickets are the execution slot's.
constexpr int AGG_WPB = 4, AGG_NGB = AGG_WPB * AGG_GPW;
int ensure_tickets(bgls_ctx* ctx, Slot* sl, size_t count);
thread_local Slot* tl_slot = nullptr;   // the execution slot of the API call running on this thread (set by SlotLock)
template <class E> int agg_blocks_per_sm(bgls_ctx* ctx, int* out) {
    static int cached = 0;
    if (!cached) {
        CU(cudaFuncSetAttribute(k_agg<E, AGG_WPB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)agg_smem_bytes<E, AGG_NGB>()));
        int b = 0;
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, k_agg<E, AGG_WPB>, AGG_WPB * 32, agg_smem_bytes<E, AGG_NGB>()));
        cached = b > 0 ? b : 1;
    }
    *out = cached;
    return BGLS_OK;
}
size_t agg_grid(size_t n, int sms, int per_sm) {
    size_t nb = (n + AGG_NGB - 1) / AGG_NGB;
    const size_t cap = std::min<size_t>((size_t)sms * per_sm, 4095);
    if (nb > cap) nb = cap;
    return nb ? nb : 1;
}
template <class E>
int enqueue_aggregate(bgls_ctx* ctx, const uint8_t* d_pts, size_t n, uint8_t* d_out, void* work, cudaStream_t s) {
    if (!tl_slot) return fail(ctx, BGLS_ERR_ARG, "aggregation called outside an execution slot");
    int per_sm = 1;
    int rc = agg_blocks_per_sm<E>(ctx, &per_sm);
    if (rc) return rc;
    if (ctx->agg_blocks_per_sm > 0 && ctx->agg_blocks_per_sm < per_sm) per_sm = ctx->agg_blocks_per_sm;
    const size_t nb = agg_grid(n, ctx->sms, per_sm);
    rc = ensure_tickets(ctx, tl_slot, agg_tree_tickets(nb, AGG_NGB) + 1);
    if (rc) return rc;
    k_agg<E, AGG_WPB><<<(unsigned)nb, AGG_WPB * 32, agg_smem_bytes<E, AGG_NGB>(), s>>>(d_pts, n, (uint32_t*)work, tl_slot->tickets, d_out, nullptr);
    ctx->launches++;
    CU(cudaGetLastError());
    return BGLS_OK;
}

// peers / fused_send: the multi-GPU exchange rides in the slot engine's launch when that engine runs (then *fused_send = true)
int pairing_dev(bgls_ctx* ctx, int curve, const void* d_g1, const void* d_g2, size_t n, int do_final, void* d_out,
                void* d_flag, void* work, cudaStream_t s, const SvPeers* peers = nullptr, bool* fused_send = nullptr);
// the same for the waiting side: the bounded wait for the peers' flags opens the slot engine's finishing launch (*fused_wait = true)
struct XchWait {
    const unsigned long long* flags;
    unsigned long long epoch;
    int* err;
};
int finish_bytes_dev(bgls_ctx* ctx, int curve, const void* d_partials, size_t k, int do_final, void* d_out, void* d_flag, void* work, cudaStream_t s,
                     const XchWait* wait = nullptr, bool* fused_wait = nullptr);
int batch_dev(bgls_ctx* ctx, int curve, const void* d_g1, const void* d_g2, const void* d_off, size_t nbatch, size_t total,
              void* d_ok, void* work, cudaStream_t s);

int pairing_dev_thread(bgls_ctx* ctx, int curve, const void* d_g1, const void* d_g2, size_t n, int do_final, void* d_out,
                void* d_flag, void* work, cudaStream_t s, bool prof) {
    if (curve == BGLS_ALTBN128)
        return enqueue_pairing<BN254>(ctx, (const uint8_t*)d_g1, (const uint8_t*)d_g2, n, do_final, (uint8_t*)d_out, (int*)d_flag, work, s, prof);
    return enqueue_pairing<BLS381>(ctx, (const uint8_t*)d_g1, (const uint8_t*)d_g2, n, do_final, (uint8_t*)d_out, (int*)d_flag, work, s, prof);
}
int finish_bytes_dev_thread(bgls_ctx* ctx, int curve, const void* d_partials, size_t k, int do_final, void* d_out, void* d_flag, cudaStream_t s) {
    if (curve == BGLS_ALTBN128) k_finish<BN254, true><<<1, TB, 0, s>>>(d_partials, k, do_final, (uint8_t*)d_out, (int*)d_flag);
    else k_finish<BLS381, true><<<1, TB, 0, s>>>(d_partials, k, do_final, (uint8_t*)d_out, (int*)d_flag);
    ctx->launches++;
    CU(cudaGetLastError());
    return BGLS_OK;
}
int aggregate_dev(bgls_ctx* ctx, int curve, int group, const void* d_pts, size_t n, void* d_out, void* work, cudaStream_t s) {
    const size_t rec = 2 * group * fp_bytes(curve);
    const uint8_t* p = (const uint8_t*)d_pts;
    uint8_t* o = (uint8_t*)d_out;
    if (curve == BGLS_ALTBN128) {
        if (group == 1) return enqueue_aggregate<AggFp<BN254>>(ctx, p, n, o, work, s);
        return enqueue_aggregate<AggFp2<BN254>>(ctx, p, n, o, work, s);
    }
    if (group == 1) return enqueue_aggregate<AggFp<BLS381>>(ctx, p, n, o, work, s);
    return enqueue_aggregate<AggFp2<BLS381>>(ctx, p, n, o, work, s);
}
int scale_dev(bgls_ctx* ctx, int curve, int group, const void* d_pts, const void* d_sc, size_t n, void* d_out, cudaStream_t s) {
    const size_t rec = 2 * group * fp_bytes(curve);
    const uint8_t *p = (const uint8_t*)d_pts, *sc = (const uint8_t*)d_sc;
    uint8_t* o = (uint8_t*)d_out;
    const int nb = (int)((n + TB - 1) / TB);
    if (n == 0) return BGLS_OK;
    if (curve == BGLS_ALTBN128) {
        if (group == 1) k_scale<BN254, Fp<BN254>><<<nb, TB, 0, s>>>(p, sc, n, rec, o);
        else k_scale<BN254, Fp2<BN254>><<<nb, TB, 0, s>>>(p, sc, n, rec, o);
    } else {
        if (group == 1) k_scale<BLS381, Fp<BLS381>><<<nb, TB, 0, s>>>(p, sc, n, rec, o);
        else k_scale<BLS381, Fp2<BLS381>><<<nb, TB, 0, s>>>(p, sc, n, rec, o);
    }
    ctx->launches++;
    CU(cudaGetLastError());
    return BGLS_OK;
}
int batch_dev_thread(bgls_ctx* ctx, int curve, const void* d_g1, const void* d_g2, const void* d_off, size_t nbatch, size_t total,
              void* d_ok, void* work, cudaStream_t s) {
    const int nb1 = (int)((total + TB - 1) / TB), nb2 = (int)((nbatch + TB - 1) / TB);
    if (curve == BGLS_ALTBN128) {
        if (total) k_miller_values<BN254><<<nb1, TB, 0, s>>>((const uint8_t*)d_g1, (const uint8_t*)d_g2, total, (Fp12<BN254>*)work);
        if (nbatch) k_batch_finish<BN254><<<nb2, TB, 0, s>>>((const Fp12<BN254>*)work, (const uint64_t*)d_off, nbatch, (uint8_t*)d_ok);
    } else {
        if (total) k_miller_values<BLS381><<<nb1, TB, 0, s>>>((const uint8_t*)d_g1, (const uint8_t*)d_g2, total, (Fp12<BLS381>*)work);
        if (nbatch) k_batch_finish<BLS381><<<nb2, TB, 0, s>>>((const Fp12<BLS381>*)work, (const uint64_t*)d_off, nbatch, (uint8_t*)d_ok);
    }
    ctx->launches += (total ? 1 : 0) + (nbatch ? 1 : 0);
    CU(cudaGetLastError());
    return BGLS_OK;
}

// ---- machine tables upload
template <class M, class MT, class F, class FT, class P, class PT> int upload_mach(bgls_ctx* ctx, MachDev& d) {
    struct Part { const void* src; size_t bytes; const void** dst; };
    const Part parts[] = {
        {MT::consts(), sizeof(uint32_t) * M::NCONST * M::L, (const void**)&d.m.consts},
        {MT::hdr(), sizeof(uint32_t) * M::NPHASE, (const void**)&d.m.hdr},
        {MT::rec(), sizeof(uint16_t) * M::NPHASE * M::LANES * M::REC, (const void**)&d.m.rec},
        {FT::consts(), sizeof(uint32_t) * F::NCONST * F::L, (const void**)&d.f.consts},
        {FT::hdr(), sizeof(uint32_t) * F::NPHASE, (const void**)&d.f.hdr},
        {FT::rec(), sizeof(uint16_t) * F::NPHASE * F::LANES * F::REC, (const void**)&d.f.rec},
        {MT::prog_MILLER(), sizeof(uint16_t) * M::MILLER_LEN, (const void**)&d.miller},
        {MT::prog_MUL_AB(), sizeof(uint16_t) * M::MUL_AB_LEN, (const void**)&d.mul_ab},
        {MT::prog_MUL_BA(), sizeof(uint16_t) * M::MUL_BA_LEN, (const void**)&d.mul_ba},
        {MT::prog_IMPORT_A(), sizeof(uint16_t) * M::IMPORT_A_LEN, (const void**)&d.import_a},
        {FT::prog_FINALEXP(), sizeof(uint16_t) * F::FINALEXP_LEN, (const void**)&d.finalexp},
        {FT::prog_EXPORT(), sizeof(uint16_t) * F::EXPORT_LEN, (const void**)&d.export_},
        {PT::consts(), sizeof(uint32_t) * P::NCONST * P::L, (const void**)&d.p.consts},
        {PT::hdr(), sizeof(uint32_t) * P::NPHASE, (const void**)&d.p.hdr},
        {PT::rec(), sizeof(uint16_t) * P::NPHASE * P::LANES * P::REC, (const void**)&d.p.rec},
        {PT::prog_MILLER(), sizeof(uint16_t) * P::MILLER_LEN, (const void**)&d.miller_p},
        {PT::prog_MULACC(), sizeof(uint16_t) * P::MULACC_LEN, (const void**)&d.mulacc_p},
    };
    size_t total = 0;
    for (const Part& p : parts) total += align_up(p.bytes);
    CU(cudaMalloc(&d.blob, total));
    size_t off = 0;
    for (const Part& p : parts) {
        CU(cudaMemcpy((char*)d.blob + off, p.src, p.bytes, cudaMemcpyHostToDevice));
        *p.dst = (char*)d.blob + off;
        off += align_up(p.bytes);
    }
    CU(cudaFuncSetAttribute(k_mach_miller<M, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mach_smem_bytes<M>(MGPB)));
    CU(cudaFuncSetAttribute(k_mach_miller<M, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mach_smem_bytes<M>(MWPB)));
    CU(cudaFuncSetAttribute(k_mach_reduce<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mach_smem_bytes<M>(MGPB)));
    CU(cudaFuncSetAttribute(k_mach_import<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mach_smem_bytes<M>(MGPB)));
    CU(cudaFuncSetAttribute(k_mach_finish<F, M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mach_smem_bytes<F>(1)));
    CU(cudaFuncSetAttribute(k_mach_finish_lean<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((size_t)F::NS * F::L * sizeof(uint32_t))));
    CU(cudaFuncSetAttribute(k_mach_finish_lean<F>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    CU(cudaFuncSetAttribute(k_mach_miller32<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mach_smem_bytes<P>(MP_MAXW)));
    CU(cudaFuncSetAttribute(k_mach_tree32<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mach_smem_bytes<P>(MT_MAXW)));
    return BGLS_OK;
}

constexpr size_t MCHUNK = 8;  // fan-in of one product-tree level
constexpr size_t MHYBRID = 16384;  // pairs: above this the thread-per-pair Miller kernel wins (measured: 0.45 vs 0.65 us/pair)
constexpr size_t MSMALL = 0;  // GPW=1 threshold: measured slower on B200 (1.55 ms vs 1.19 ms at 1025 pairs: the IMAD pipe is
                              // charged per warp instruction, so half-empty warps double the pipe work); kept for experiments

template <class M> struct CurveOf;
template <> struct CurveOf<mtab::BN254_M> { using type = BN254; };
template <> struct CurveOf<mtab::BLS381_M> { using type = BLS381; };
constexpr size_t MBATCH_SHARED_MIN = 8192;   // total pairs of a batch from which the shared-accumulator thread kernel is used

template <class M> struct PFile;
template <> struct PFile<mtab::BN254_M> { using type = mtab::BN254_MP; };
template <> struct PFile<mtab::BLS381_M> { using type = mtab::BLS381_MP; };

// Miller loops of n pairs.  fuse: the values are multiplied per block inside the kernel (product pipelines);
// returns the number of values written to `vals` (n without fusion).
template <class M>
size_t launch_miller(bgls_ctx* ctx, const MachDev& d, const uint8_t* d_g1, const uint8_t* d_g2, size_t n, uint32_t* vals, cudaStream_t s,
                     bool fuse = false) {
    if (n == 0) return 0;
    if (!ctx->miller16) {
        // one pairing per warp; the block size spreads the pairs over all SMs (small products) up to 16 warps per SM
        using P = typename PFile<M>::type;
        size_t wpb = (n + ctx->sms - 1) / ctx->sms;
        // at least 8 pairings per block: a lone launch is latency bound per warp and does not care, but with several
        // products in flight three resident blocks then give 24 instead of 21 warps per SM (+2.6 % pairings/s device
        // resident, +3.8 % end to end at 1025 pairs, same latency; profiles/r1_bh_*).  BGLS_MIN_WPB overrides.
        const size_t min_wpb = n < (size_t)ctx->min_wpb ? n : (size_t)ctx->min_wpb;
        if (wpb < min_wpb) wpb = min_wpb;
        // large products: two resident blocks per SM (shared memory: tables + one pooled slot file per warp;
        // registers: 81 / 98 per thread) give more warps per sub-partition than one block of 16
        const size_t big_wpb = P::L <= 10 ? 12 : 10;
        if (wpb > big_wpb) wpb = big_wpb;
        const size_t nb = (n + wpb - 1) / wpb;
        k_mach_miller32<P><<<(unsigned)nb, (unsigned)wpb * 32, mach_smem_bytes<P>((int)wpb), s>>>(d.p, d.miller_p, P::MILLER_LEN, d.mulacc_p, fuse ? 1 : 0,
                                                                                             d_g1, d_g2, n, vals);
        ctx->launches++;
        return fuse ? nb : n;
    }
    if (n <= MSMALL)
        k_mach_miller<M, 1><<<(unsigned)((n + MWPB - 1) / MWPB), MWPB * 32, mach_smem_bytes<M>(MWPB), s>>>(d.m, d.miller, M::MILLER_LEN, d_g1, d_g2, n, vals);
    else
        k_mach_miller<M, 2><<<(unsigned)((n + MGPB - 1) / MGPB), MWPB * 32, mach_smem_bytes<M>(MGPB), s>>>(d.m, d.miller, M::MILLER_LEN, d_g1, d_g2, n, vals);
    ctx->launches++;
    return n;
}
template <class M> size_t mach_val_bytes() { return (size_t)12 * M::L * sizeof(uint32_t); }
template <class M> size_t mach_work_bytes(size_t n) { return align_up((n + 2) * mach_val_bytes<M>()) + align_up((n / MCHUNK + 3) * mach_val_bytes<M>()); }

// product tree over `cnt` machine values in buf0 (scratch buf1); returns the buffer holding the single result
template <class M>
int mach_tree(bgls_ctx* ctx, const MachDev& d, uint32_t* buf0, uint32_t* buf1, size_t cnt, uint32_t** result, cudaStream_t s) {
    uint32_t *cur = buf0, *oth = buf1;
    if (cnt == 0) {  // empty product = 1
        k_mach_reduce<M><<<1, MWPB * 32, mach_smem_bytes<M>(MGPB), s>>>(d.m, d.mul_ab, d.mul_ba, cur, 0, MCHUNK, nullptr, 1, oth);
        ctx->launches++;
        std::swap(cur, oth);
        cnt = 1;
    }
    while (cnt > 1 && !ctx->miller16) {   // binary in-block trees, 16 values per block and level
        using P = typename PFile<M>::type;
        const size_t wpb = cnt < (size_t)MT_MAXW ? cnt : (size_t)MT_MAXW;
        const size_t nout = (cnt + wpb - 1) / wpb;
        k_mach_tree32<P><<<(unsigned)nout, (unsigned)wpb * 32, mach_smem_bytes<P>((int)wpb), s>>>(d.p, d.mulacc_p, cur, cnt, oth);
        ctx->launches++;
        std::swap(cur, oth);
        cnt = nout;
    }
    while (cnt > 1) {
        const size_t nout = (cnt + MCHUNK - 1) / MCHUNK;
        k_mach_reduce<M><<<(unsigned)((nout + MGPB - 1) / MGPB), MWPB * 32, mach_smem_bytes<M>(MGPB), s>>>(d.m, d.mul_ab, d.mul_ba, cur, cnt, MCHUNK, nullptr, nout, oth);
        ctx->launches++;
        std::swap(cur, oth);
        cnt = nout;
    }
    *result = cur;
    CU(cudaGetLastError());
    return BGLS_OK;
}

template <class M, class F>
int mach_pairing(bgls_ctx* ctx, const MachDev& d, const uint8_t* d_g1, const uint8_t* d_g2, size_t n, int do_final,
                 uint8_t* d_out, int* d_flag, void* work, cudaStream_t s) {
    uint32_t* buf0 = (uint32_t*)work;
    uint32_t* buf1 = (uint32_t*)((char*)work + align_up((n + 2) * mach_val_bytes<M>()));
    if (ctx->profiling) cudaEventRecord(ctx->ev[0], s);
    const size_t cnt = launch_miller<M>(ctx, d, d_g1, d_g2, n, buf0, s, true);
    if (ctx->profiling) cudaEventRecord(ctx->ev[1], s);
    uint32_t* res;
    int rc = mach_tree<M>(ctx, d, buf0, buf1, cnt, &res, s);
    if (rc) return rc;
    k_mach_finish<F, M><<<1, MFIN_THREADS, mach_smem_bytes<F>(1), s>>>(d.f, do_final ? d.finalexp : d.export_, do_final ? F::FINALEXP_LEN : F::EXPORT_LEN,
                                                           res, 1, d_out, d_flag, nullptr);
    ctx->launches++;
    if (ctx->profiling) cudaEventRecord(ctx->ev[2], s);
    CU(cudaGetLastError());
    return BGLS_OK;
}
// k partial products given as wire bytes -> product -> final exponentiation
template <class M, class F>
int mach_finish_bytes(bgls_ctx* ctx, const MachDev& d, const uint8_t* d_partials, size_t k, int do_final, uint8_t* d_out, int* d_flag,
                      void* work, cudaStream_t s) {
    uint32_t* buf0 = (uint32_t*)work;
    uint32_t* buf1 = (uint32_t*)((char*)work + align_up((k + 2) * mach_val_bytes<M>()));
    if (k > 0) {
        k_mach_import<M><<<(unsigned)((k + MGPB - 1) / MGPB), MWPB * 32, mach_smem_bytes<M>(MGPB), s>>>(d.m, d.import_a, M::IMPORT_A_LEN, d_partials, k, buf0);
        ctx->launches++;
    }
    uint32_t* res;
    int rc = mach_tree<M>(ctx, d, buf0, buf1, k, &res, s);
    if (rc) return rc;
    k_mach_finish<F, M><<<1, MFIN_THREADS, mach_smem_bytes<F>(1), s>>>(d.f, do_final ? d.finalexp : d.export_, do_final ? F::FINALEXP_LEN : F::EXPORT_LEN,
                                                           res, 1, d_out, d_flag, nullptr);
    ctx->launches++;
    CU(cudaGetLastError());
    return BGLS_OK;
}
// nbatch independent products (segments given by device offsets) -> identity flags
template <class M, class F>
int mach_batch(bgls_ctx* ctx, const MachDev& d, const uint8_t* d_g1, const uint8_t* d_g2, const unsigned long long* d_off, size_t nbatch,
               size_t total, uint8_t* d_ok, void* work, cudaStream_t s) {
    uint32_t* buf0 = (uint32_t*)work;
    uint32_t* buf1 = (uint32_t*)((char*)work + align_up((total + 2) * mach_val_bytes<M>()));
    if (total >= MBATCH_SHARED_MIN && total >= 16 * nbatch && !ctx->no_shared && !ctx->machine_only) {
        // throughput regime: saturated-limb thread kernel with a shared Miller accumulator per thread, one block per
        // check -> raw products as wire bytes -> machine form -> one final exponentiation (one warp) per check
        using C = typename CurveOf<M>::type;
        uint8_t* raw = (uint8_t*)buf1;
        k_miller_batch_shared<C><<<(unsigned)nbatch, TBB, 0, s>>>(d_g1, d_g2, d_off, nbatch, raw);
        k_mach_import<M><<<(unsigned)((nbatch + MGPB - 1) / MGPB), MWPB * 32, mach_smem_bytes<M>(MGPB), s>>>(d.m, d.import_a, M::IMPORT_A_LEN, raw, nbatch, buf0);
        k_mach_finish<F, M><<<(unsigned)nbatch, MFIN_THREADS, mach_smem_bytes<F>(1), s>>>(d.f, d.finalexp, F::FINALEXP_LEN, buf0, nbatch, nullptr, nullptr, d_ok);
        ctx->launches += 3;
        CU(cudaGetLastError());
        return BGLS_OK;
    }
    launch_miller<M>(ctx, d, d_g1, d_g2, total, buf0, s);
    k_mach_reduce<M><<<(unsigned)((nbatch + MGPB - 1) / MGPB), MWPB * 32, mach_smem_bytes<M>(MGPB), s>>>(d.m, d.mul_ab, d.mul_ba, buf0, total, 0, d_off, nbatch, buf1);
    k_mach_finish<F, M><<<(unsigned)nbatch, MFIN_THREADS, mach_smem_bytes<F>(1), s>>>(d.f, d.finalexp, F::FINALEXP_LEN, buf1, nbatch, nullptr, nullptr, d_ok);
    ctx->launches += 2;
    CU(cudaGetLastError());
    return BGLS_OK;
}

// ---- slot engine (slotvm.cuh): Miller loops on saturated limbs, G lanes per pair
// 8 lanes per group of 2 pairs that share one Miller accumulator: the best throughput of the shapes measured on B200
// (profiles/r2_l_*: altbn128 4.1 M pairings/s with the GPU full against 3.7 for 4 lanes per pair, 3.6 for 8 lanes per pair)
using SlotBN = svt::BN254_G8K2;
using SlotBLS = svt::BLS381_G8K2;
#ifndef BGLS_SLOT_WPB
#define BGLS_SLOT_WPB 1
#endif
constexpr int SLOT_WPB = BGLS_SLOT_WPB;   // one warp per block: 4 groups = 8 pairs, 14 / 21 KB of shared memory -> 15 / 8 resident warps per SM
constexpr size_t SLOT_PAIRS_PER_BLOCK = (size_t)SLOT_WPB * 32 / SlotBN::G * SlotBN::K;
static_assert(SlotBN::G == SlotBLS::G && SlotBN::K == SlotBLS::K, "one block shape for both curves");
template <class C, class T, class F> int upload_slot(bgls_ctx* ctx, SlotEngDev& d) {
    const size_t b0 = align_up((size_t)T::NWORDS * 4), b1 = align_up((size_t)(T::NPROG + 1) * 4), b2 = align_up((size_t)T::SEQ_LEN),
                 b3 = align_up((size_t)T::NCONST * 2 * C::N * 4), b4 = align_up((size_t)C::N * 4);
    CU(cudaMalloc(&d.blob, b0 + b1 + b2 + b3 + b4));
    char* p = (char*)d.blob;
    CU(cudaMemcpy(p, T::code(), (size_t)T::NWORDS * 4, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(p + b0, T::offsets(), (size_t)(T::NPROG + 1) * 4, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(p + b0 + b1, T::sequence(), (size_t)T::SEQ_LEN, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(p + b0 + b1 + b2, T::consts(), (size_t)T::NCONST * 2 * C::N * 4, cudaMemcpyHostToDevice));
    d.tb.code = (const uint32_t*)p;
    d.tb.offs = (const uint32_t*)(p + b0);
    d.tb.seq = (const uint8_t*)(p + b0 + b1);
    d.tb.consts = (const uint32_t*)(p + b0 + b1 + b2);
    CU(cudaMemcpy(p + b0 + b1 + b2 + b3, T::mach_r(), (size_t)C::N * 4, cudaMemcpyHostToDevice));
    d.tb.mach_r = (const uint32_t*)(p + b0 + b1 + b2 + b3);
    constexpr int NPB = SLOT_WPB * 32 / T::G;
    using FIN = MachFinisher<F>;
    CU(cudaFuncSetAttribute(k_slot_miller<C, T, SLOT_WPB, FIN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            (int)std::max(sv_smem_bytes<C, T, NPB>(), FIN::SMEM_BYTES)));
    // every kernel of the throughput pipeline asks for the same L1 / shared-memory split: kernels that prefer different
    // carve-outs cannot share an SM, and an SM drains before it switches (measured: the pipeline was capped at 2.6 M
    // pairings/s however many products were in flight, the Miller kernel alone reached 3.5 M)
    CU(cudaFuncSetAttribute(k_slot_miller<C, T, SLOT_WPB, FIN>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    CU(cudaFuncSetAttribute(k_slot_miller<C, T, SLOT_WPB, SvNoFinish>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sv_smem_bytes<C, T, NPB>()));
    CU(cudaFuncSetAttribute(k_slot_miller<C, T, SLOT_WPB, SvNoFinish>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    CU(cudaFuncSetAttribute(k_slot_miller<C, T, SLOT_WPB, SvWireFinisher<C, T>>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sv_smem_bytes<C, T, NPB>()));
    CU(cudaFuncSetAttribute(k_slot_miller<C, T, SLOT_WPB, SvWireFinisher<C, T>>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    CU(cudaFuncSetAttribute(k_mach_finish_lean<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FIN::SMEM_BYTES));
    CU(cudaFuncSetAttribute(k_mach_finish_lean<F>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    CU(cudaFuncSetAttribute(k_slot_finish_bytes<C, T, SLOT_WPB, FIN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            (int)std::max(sv_smem_bytes<C, T, NPB>(), FIN::SMEM_BYTES)));
    CU(cudaFuncSetAttribute(k_slot_finish_bytes<C, T, SLOT_WPB, FIN>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    return BGLS_OK;
}
size_t slot_blocks(size_t n) { return (n + SLOT_PAIRS_PER_BLOCK - 1) / SLOT_PAIRS_PER_BLOCK; }
constexpr size_t SLOT_FANIN = 2 * (SLOT_WPB * 32 / SlotBN::G);   // values multiplied by one block of the in-launch product tree
// scratch of the slot pipeline for n pairs: [tree levels | one machine-form value]
size_t slot_work_bytes(int curve, size_t n) {
    const size_t N = curve == BGLS_ALTBN128 ? 8 : 12, L = curve == BGLS_ALTBN128 ? 10 : 14;
    const size_t nb = slot_blocks(n);
    return align_up((sv_tree_words(nb, SLOT_FANIN, N) + 1) * 4) + align_up(12 * L * 4);
}
// per-slot tickets of the in-launch product tree: zero between launches (the last arriver of a group resets its ticket)
int ensure_tickets(bgls_ctx* ctx, Slot* sl, size_t count) {
    if (count <= sl->tickets_cap) return BGLS_OK;
    CU(cudaDeviceSynchronize());   // growth only (first call of a size): see ensure_scratch
    if (sl->tickets) CU(cudaFree(sl->tickets));
    sl->tickets = nullptr;
    sl->tickets_cap = 0;
    const size_t want = std::max<size_t>(4096, count * 2);
    CU(cudaMalloc((void**)&sl->tickets, want * 4));
    CU(cudaMemset(sl->tickets, 0, want * 4));
    sl->tickets_cap = want;
    return BGLS_OK;
}
// Miller loops of n pairs, their product tree and the machine's final exponentiation (or plain export) in ONE launch:
// k_slot_miller with the low-footprint finisher as the tail of the block that ends up with the product.
template <class C, class T, class F>
int slot_pairing(bgls_ctx* ctx, const SlotEngDev& se, const MachDev& md, const uint8_t* d_g1, const uint8_t* d_g2, size_t n, int do_final,
                 uint8_t* d_out, int* d_flag, void* work, cudaStream_t s, const SvPeers* peers = nullptr) {
    using FIN = MachFinisher<F>;
    constexpr int NPB = SLOT_WPB * 32 / T::G;
    constexpr size_t smem = std::max(sv_smem_bytes<C, T, NPB>(), FIN::SMEM_BYTES);
    static_assert(T::MACH_L == F::L, "machine limb count");
    static_assert(SLOT_FANIN == 2 * NPB, "fan-in of the in-launch tree");
    if (!tl_slot) return fail(ctx, BGLS_ERR_ARG, "slot engine called outside an execution slot");
    const size_t nb = slot_blocks(n);
    int rc = ensure_tickets(ctx, tl_slot, sv_tree_counters(nb, SLOT_FANIN) + 1);
    if (rc) return rc;
    uint32_t* levels = (uint32_t*)work;
    uint32_t* mval = (uint32_t*)((char*)levels + align_up((sv_tree_words(nb, SLOT_FANIN, C::N) + 1) * 4));
    MachFinishArgs fa{md.f, do_final ? md.finalexp : md.export_, do_final ? F::FINALEXP_LEN : F::EXPORT_LEN, d_out, d_flag, nullptr};
    if (ctx->profiling) cudaEventRecord(ctx->ev[0], s);
    if (!do_final) {
        // raw Miller product (a shard of a multi-GPU verification): wire export by the slot engine itself, so the blocks carry
        // no finisher footprint; the stores into the peers' mailboxes follow in the same launch when `peers` is given
        using WF = SvWireFinisher<C, T>;
        k_slot_miller<C, T, SLOT_WPB, WF><<<(unsigned)nb, SLOT_WPB * 32, sv_smem_bytes<C, T, NPB>(), s>>>(
            se.tb, d_g1, d_g2, n, levels, tl_slot->tickets, mval, 0, typename WF::Args{d_out}, SvBatch{}, ctx->trace, peers ? *peers : SvPeers{});
        ctx->launches++;
        if (ctx->profiling) { cudaEventRecord(ctx->ev[1], s); cudaEventRecord(ctx->ev[2], s); }
        CU(cudaGetLastError());
        return BGLS_OK;
    }
    if ((ctx->slot_split < 0 ? C::IS_BN : ctx->slot_split == 1) && !peers) {
        // Miller blocks without the finisher's shared-memory footprint (altbn128: 14 KB instead of 19 KB per one-warp block,
        // 15 instead of 11 resident blocks per SM), final exponentiation as a second one-warp launch: +5 % on altbn128,
        // -10 % on bls12-381 where the longer exponentiation is better hidden inside the launch
        k_slot_miller<C, T, SLOT_WPB, SvNoFinish><<<(unsigned)nb, SLOT_WPB * 32, sv_smem_bytes<C, T, NPB>(), s>>>(
            se.tb, d_g1, d_g2, n, levels, tl_slot->tickets, mval, T::MACH_L, SvNoFinish::Args{0}, SvBatch{}, ctx->trace, SvPeers{});
        if (ctx->profiling) cudaEventRecord(ctx->ev[1], s);
        k_mach_finish_lean<F><<<1, 32, FIN::SMEM_BYTES, s>>>(fa, mval);
        ctx->launches += 2;
        if (ctx->profiling) cudaEventRecord(ctx->ev[2], s);
        CU(cudaGetLastError());
        return BGLS_OK;
    }
    k_slot_miller<C, T, SLOT_WPB, FIN><<<(unsigned)nb, SLOT_WPB * 32, smem, s>>>(se.tb, d_g1, d_g2, n, levels, tl_slot->tickets, mval, T::MACH_L, fa, SvBatch{}, ctx->trace,
                                                                                 peers ? *peers : SvPeers{});
    ctx->launches++;
    if (ctx->profiling) { cudaEventRecord(ctx->ev[1], s); cudaEventRecord(ctx->ev[2], s); }
    CU(cudaGetLastError());
    return BGLS_OK;
}
// product of k wire-form Miller values (k <= SLOT_FANIN) + final exponentiation / export in one launch
template <class C, class T, class F>
int slot_finish_bytes(bgls_ctx* ctx, const SlotEngDev& se, const MachDev& md, const uint8_t* d_partials, size_t k, int do_final, uint8_t* d_out,
                      int* d_flag, void* work, cudaStream_t s, const XchWait* wait = nullptr) {
    using FIN = MachFinisher<F>;
    constexpr int NPB = SLOT_WPB * 32 / T::G;
    constexpr size_t smem = std::max(sv_smem_bytes<C, T, NPB>(), FIN::SMEM_BYTES);
    MachFinishArgs fa{md.f, do_final ? md.finalexp : md.export_, do_final ? F::FINALEXP_LEN : F::EXPORT_LEN, d_out, d_flag, nullptr};
    k_slot_finish_bytes<C, T, SLOT_WPB, FIN><<<1, SLOT_WPB * 32, smem, s>>>(se.tb, d_partials, (int)k, (uint32_t*)work, T::MACH_L, fa,
                                                                            wait ? wait->flags : nullptr, wait ? wait->epoch : 0ull, wait ? wait->err : nullptr);
    ctx->launches++;
    CU(cudaGetLastError());
    return BGLS_OK;
}
// nbatch independent products in one launch (k_slot_plan + k_slot_miller in batch form): verdict bytes in d_ok
size_t slot_batch_blocks(size_t nbatch, size_t total) { return total / SLOT_PAIRS_PER_BLOCK + nbatch; }   // upper bound
size_t slot_batch_work_bytes(int curve, size_t nbatch, size_t total) {
    const size_t N = curve == BGLS_ALTBN128 ? 8 : 12, L = curve == BGLS_ALTBN128 ? 10 : 14, nb = slot_batch_blocks(nbatch, total);
    return 3 * align_up((nbatch + 1) * 4) + align_up((nb + nb / 4 + 8 * nbatch) * 12 * N * 4) + align_up(nbatch * 12 * L * 4);
}
template <class C, class T, class F>
int slot_batch(bgls_ctx* ctx, const SlotEngDev& se, const MachDev& md, const uint8_t* d_g1, const uint8_t* d_g2, const unsigned long long* d_off,
               size_t nbatch, size_t total, uint8_t* d_ok, void* work, cudaStream_t s) {
    using FIN = MachFinisher<F>;
    constexpr int NPB = SLOT_WPB * 32 / T::G;
    constexpr size_t smem = std::max(sv_smem_bytes<C, T, NPB>(), FIN::SMEM_BYTES);
    if (!tl_slot) return fail(ctx, BGLS_ERR_ARG, "slot engine called outside an execution slot");
    const size_t nb = slot_batch_blocks(nbatch, total);
    int rc = ensure_tickets(ctx, tl_slot, nb / 4 + 8 * nbatch + 8);   // >= sum of sv_tree_counters over the products
    if (rc) return rc;
    char* w = (char*)work;
    unsigned* bstart = (unsigned*)w;
    unsigned* lstart = (unsigned*)(w + align_up((nbatch + 1) * 4));
    unsigned* tstart = (unsigned*)(w + 2 * align_up((nbatch + 1) * 4));
    uint32_t* levels = (uint32_t*)(w + 3 * align_up((nbatch + 1) * 4));
    uint32_t* mvals = (uint32_t*)((char*)levels + align_up((nb + nb / 4 + 8 * nbatch) * 12 * C::N * 4));
    k_slot_plan<<<1, 1024, 0, s>>>(d_off, nbatch, (unsigned)SLOT_PAIRS_PER_BLOCK, (unsigned)SLOT_FANIN, bstart, lstart, tstart);
    MachFinishArgs fa{md.f, md.finalexp, F::FINALEXP_LEN, nullptr, nullptr, nullptr};
    SvBatch b{d_off, bstart, lstart, tstart, nbatch, d_ok};
    k_slot_miller<C, T, SLOT_WPB, FIN><<<(unsigned)nb, SLOT_WPB * 32, smem, s>>>(se.tb, d_g1, d_g2, 0, levels, tl_slot->tickets, mvals, T::MACH_L, fa, b, ctx->trace, SvPeers{});
    ctx->launches += 2;
    CU(cudaGetLastError());
    return BGLS_OK;
}

// load estimate: pairing calls inside the library (host-buffer entry points) or recently enqueued (device-resident ones)
long long now_us() { return std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int busy_estimate(bgls_ctx* ctx) {
    const long long t = now_us();
    int recent = 0;
    for (auto& sl : ctx->slots) recent += (t - sl.last_use_us.load(std::memory_order_relaxed)) < 3000 ? 1 : 0;
    const int hc = ctx->host_calls.load(std::memory_order_relaxed);
    return hc > recent ? hc : recent;
}
constexpr size_t SLOT_MIN_PAIRS = 32;     // below this a product is a handful of warps either way: the machine's one warp per pairing is faster
constexpr size_t SLOT_BIG_PAIRS = 2048;   // from here one product alone fills the GPU with the slot engine
constexpr int SLOT_MIN_BUSY = 3;          // products in flight from which throughput matters more than the latency of one
bool use_slot_engine(bgls_ctx* ctx, size_t n) {
    if (ctx->engine == ENGINE_SLOT) return n > 0;
    if (ctx->engine != ENGINE_AUTO || n < SLOT_MIN_PAIRS) return false;
    return n >= SLOT_BIG_PAIRS || busy_estimate(ctx) >= SLOT_MIN_BUSY;
}

size_t mach_work_for(int curve, size_t n) {
    return curve == BGLS_ALTBN128 ? mach_work_bytes<mtab::BN254_M>(n) : mach_work_bytes<mtab::BLS381_M>(n);
}
int pairing_dev(bgls_ctx* ctx, int curve, const void* d_g1, const void* d_g2, size_t n, int do_final, void* d_out,
                void* d_flag, void* work, cudaStream_t s, const SvPeers* peers, bool* fused_send) {
    if (fused_send) *fused_send = false;
    if (ctx->thread_engine) return pairing_dev_thread(ctx, curve, d_g1, d_g2, n, do_final, d_out, d_flag, work, s, ctx->profiling);
    if (use_slot_engine(ctx, n)) {
        // throughput regime: Miller loops and product tree on the slot engine (saturated limbs, 8 lanes per group of 2 pairs
        // that share their accumulator), then the machine's final exponentiation in its low-footprint form
        if (fused_send && peers && !do_final) *fused_send = true;
        const SvPeers* pp = (fused_send && *fused_send) ? peers : nullptr;
        if (curve == BGLS_ALTBN128)
            return slot_pairing<BN254, SlotBN, mtab::BN254_F>(ctx, ctx->sloteng[0], ctx->mach[0], (const uint8_t*)d_g1, (const uint8_t*)d_g2,
                                                                              n, do_final, (uint8_t*)d_out, (int*)d_flag, work, s, pp);
        return slot_pairing<BLS381, SlotBLS, mtab::BLS381_F>(ctx, ctx->sloteng[1], ctx->mach[1], (const uint8_t*)d_g1, (const uint8_t*)d_g2,
                                                                              n, do_final, (uint8_t*)d_out, (int*)d_flag, work, s, pp);
    }
    if (n >= MHYBRID && !ctx->machine_only) {
        // throughput regime: the thread-per-pair Miller kernel (Karatsuba towers, saturated limbs) does ~1.45x more
        // pairs/s than the machine once the GPU is full; its raw product is handed to the machine for the
        // latency-critical final exponentiation.  Layout of `work`: [thread partials | raw bytes | machine scratch]
        const size_t F = fp_bytes(curve);
        uint8_t* tw = (uint8_t*)work;
        uint8_t* raw = tw + align_up((size_t)blocks_for(n) * fp12_dev_bytes(curve));
        uint8_t* mw = raw + align_up(12 * F);
        if (ctx->profiling) cudaEventRecord(ctx->ev[0], s);
        int rc = pairing_dev_thread(ctx, curve, d_g1, d_g2, n, 0, raw, nullptr, tw, s, false);   // no shared state is toggled
        if (rc) return rc;
        if (ctx->profiling) cudaEventRecord(ctx->ev[1], s);
        rc = finish_bytes_dev(ctx, curve, raw, 1, do_final, d_out, d_flag, mw, s);
        if (ctx->profiling) cudaEventRecord(ctx->ev[2], s);
        return rc;
    }
    if (curve == BGLS_ALTBN128)
        return mach_pairing<mtab::BN254_M, mtab::BN254_F>(ctx, ctx->mach[0], (const uint8_t*)d_g1, (const uint8_t*)d_g2, n, do_final, (uint8_t*)d_out, (int*)d_flag, work, s);
    return mach_pairing<mtab::BLS381_M, mtab::BLS381_F>(ctx, ctx->mach[1], (const uint8_t*)d_g1, (const uint8_t*)d_g2, n, do_final, (uint8_t*)d_out, (int*)d_flag, work, s);
}
int finish_bytes_dev(bgls_ctx* ctx, int curve, const void* d_partials, size_t k, int do_final, void* d_out, void* d_flag, void* work, cudaStream_t s,
                     const XchWait* wait, bool* fused_wait) {
    if (fused_wait) *fused_wait = false;
    if (ctx->thread_engine) return finish_bytes_dev_thread(ctx, curve, d_partials, k, do_final, d_out, d_flag, s);
    // throughput regime (several verifications in flight): one launch on the slot engine instead of import + tree + finish
    // (a caller that passes `wait` has already decided for this path and launched no wait kernel)
    if (k <= SLOT_FANIN && (wait || ctx->engine == ENGINE_SLOT || (ctx->engine == ENGINE_AUTO && busy_estimate(ctx) >= SLOT_MIN_BUSY))) {
        const XchWait* w = (wait && fused_wait && d_flag) ? wait : nullptr;
        if (w) *fused_wait = true;
        if (curve == BGLS_ALTBN128)
            return slot_finish_bytes<BN254, SlotBN, mtab::BN254_F>(ctx, ctx->sloteng[0], ctx->mach[0], (const uint8_t*)d_partials, k, do_final,
                                                                   (uint8_t*)d_out, (int*)d_flag, work, s, w);
        return slot_finish_bytes<BLS381, SlotBLS, mtab::BLS381_F>(ctx, ctx->sloteng[1], ctx->mach[1], (const uint8_t*)d_partials, k, do_final,
                                                                  (uint8_t*)d_out, (int*)d_flag, work, s, w);
    }
    if (curve == BGLS_ALTBN128)
        return mach_finish_bytes<mtab::BN254_M, mtab::BN254_F>(ctx, ctx->mach[0], (const uint8_t*)d_partials, k, do_final, (uint8_t*)d_out, (int*)d_flag, work, s);
    return mach_finish_bytes<mtab::BLS381_M, mtab::BLS381_F>(ctx, ctx->mach[1], (const uint8_t*)d_partials, k, do_final, (uint8_t*)d_out, (int*)d_flag, work, s);
}
constexpr size_t SLOT_BATCH_MIN_PAIRS = 2048;   // total pairs from which a batch fills the GPU with the slot engine
int batch_dev(bgls_ctx* ctx, int curve, const void* d_g1, const void* d_g2, const void* d_off, size_t nbatch, size_t total,
              void* d_ok, void* work, cudaStream_t s) {
    if (ctx->thread_engine) return batch_dev_thread(ctx, curve, d_g1, d_g2, d_off, nbatch, total, d_ok, work, s);
    if (nbatch && (ctx->engine == ENGINE_SLOT || (ctx->engine == ENGINE_AUTO && total >= SLOT_BATCH_MIN_PAIRS))) {
        if (curve == BGLS_ALTBN128)
            return slot_batch<BN254, SlotBN, mtab::BN254_F>(ctx, ctx->sloteng[0], ctx->mach[0], (const uint8_t*)d_g1, (const uint8_t*)d_g2,
                                                            (const unsigned long long*)d_off, nbatch, total, (uint8_t*)d_ok, work, s);
        return slot_batch<BLS381, SlotBLS, mtab::BLS381_F>(ctx, ctx->sloteng[1], ctx->mach[1], (const uint8_t*)d_g1, (const uint8_t*)d_g2,
                                                            (const unsigned long long*)d_off, nbatch, total, (uint8_t*)d_ok, work, s);
    }
    if (curve == BGLS_ALTBN128)
        return mach_batch<mtab::BN254_M, mtab::BN254_F>(ctx, ctx->mach[0], (const uint8_t*)d_g1, (const uint8_t*)d_g2, (const unsigned long long*)d_off, nbatch, total, (uint8_t*)d_ok, work, s);
    return mach_batch<mtab::BLS381_M, mtab::BLS381_F>(ctx, ctx->mach[1], (const uint8_t*)d_g1, (const uint8_t*)d_g2, (const unsigned long long*)d_off, nbatch, total, (uint8_t*)d_ok, work, s);
}
// scratch needed by the pairing pipelines for n pairs (either engine)
size_t pairing_work_bytes(bgls_ctx* ctx, int curve, size_t n) {
    const size_t t = align_up((size_t)blocks_for(n) * fp12_dev_bytes(curve));
    const size_t m = mach_work_for(curve, n);
    if (ctx->thread_engine) return t;
    // the engine is chosen per call (load dependent): the scratch must fit whichever runs
    const size_t sv = slot_work_bytes(curve, n);
    size_t w = m > t ? m : t;
    if (n >= MHYBRID && !ctx->machine_only) w = std::max(w, t + align_up(12 * fp_bytes(curve)) + mach_work_for(curve, 1));
    return std::max(w, sv);
}

// Acquires an execution slot for the duration of one API call.
//   host-buffer calls: any free slot (round robin when all are busy), work runs on the slot's own stream;
//   device-resident calls: the slot ordered on the caller's stream; a slot taken over from another stream first
//   waits for that stream, so scratch is never shared by work that is not stream-ordered.
struct SlotLock {
    Slot* s = nullptr;
    cudaStream_t user = nullptr;
    bool dev = false;
    // host-buffer call: any free slot, work runs on the slot's own stream
    explicit SlotLock(bgls_ctx* c) {
        // the highest free slot that no device-resident stream owns (those are handed out from index 0 up): a caller that
        // makes one call after another keeps getting the same slot, whose scratch has already grown to its sizes; C
        // concurrent callers settle on C slots.  (Round robin made each of the first NSLOT calls pay a scratch allocation.)
        for (int pass = 0; pass < 2 && !s; pass++)
            for (int i = NSLOT - 1; i >= 0 && !s; i--) {
                Slot* t = &c->slots[i];
                if (pass == 0 && t->owned) continue;   // unlocked read: a hint only
                if (t->mu.try_lock()) s = t;
            }
        if (!s) { s = &c->slots[NSLOT - 1 - c->rr.fetch_add(1) % NSLOT]; s->mu.lock(); }
        bool had = false;
        {
            std::lock_guard<std::mutex> lk(c->own_mu);
            had = s->owned;
            s->owned = false;
            s->owner = nullptr;
        }
        // the scratch may still be in use by device-resident work: order this call's stream after the event that the last
        // such call recorded (no stream handle of the caller is kept or touched)
        if (had) cudaStreamWaitEvent(s->stream, s->ev_done, 0);
        tl_slot = s;
    }
    // device-resident call: the slot ordered on the caller's stream; a slot taken over from other work first waits (on the
    // device) for the event recorded at the end of that work, so scratch is never shared by work that is not ordered
    SlotLock(bgls_ctx* c, cudaStream_t user_) : user(user_), dev(true) {
        {
            std::lock_guard<std::mutex> lk(c->own_mu);
            for (int i = 0; i < NSLOT && !s; i++)
                if (c->slots[i].owned && c->slots[i].owner == user) s = &c->slots[i];
            for (int i = 0; i < NSLOT && !s; i++)
                if (!c->slots[i].owned) s = &c->slots[i];
            if (!s) s = &c->slots[c->rr.fetch_add(1) % NSLOT];
        }
        s->mu.lock();
        cudaStream_t prev = nullptr;
        bool had = false;
        {
            std::lock_guard<std::mutex> lk(c->own_mu);
            had = s->owned;
            prev = s->owner;
            s->owned = true;
            s->owner = user;
        }
        if (had && prev != user) cudaStreamWaitEvent(user, s->ev_done, 0);
        else if (!had) {
            // last used by a host-buffer call (synchronous: its stream is idle) or never
            cudaEventRecord(s->ev_done, s->stream);
            cudaStreamWaitEvent(user, s->ev_done, 0);
        }
        s->last_use_us.store(now_us(), std::memory_order_relaxed);
        tl_slot = s;
    }
    ~SlotLock() {
        if (dev) cudaEventRecord(s->ev_done, user);   // everything this call enqueued precedes the event
        tl_slot = nullptr;
        s->mu.unlock();
    }
    SlotLock(const SlotLock&) = delete;
    SlotLock& operator=(const SlotLock&) = delete;
};

}  // namespace

extern "C" {

const char* bgls_version(void) { return kVersion; }

int bgls_ctx_create(int device, bgls_ctx** out) {
    if (!out) return BGLS_ERR_ARG;
    *out = nullptr;
    // Verifications in flight live on their own streams; CUDA maps streams onto 8 hardware queues by default and a
    // queue entry that waits (a copy behind a kernel) blocks the entries of the other streams behind it.  32 queues keep
    // up to 32 calls independent (measured: 2.1-2.8 M -> 3.1-3.4 M pairings/s with 16-32 host threads).  Only effective
    // when this is the first CUDA call of the process; a host that initialises CUDA itself sets the variable itself.
    setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0);
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) return BGLS_ERR_NODEV;
    bgls_ctx* ctx = new bgls_ctx();
    ctx->device = device;
    bool ok = cudaSetDevice(device) == cudaSuccess;
    for (int i = 0; ok && i < NSLOT; i++) {
        Slot& sl = ctx->slots[i];
        ok = cudaStreamCreateWithFlags(&sl.stream, cudaStreamNonBlocking) == cudaSuccess &&
             cudaStreamCreateWithFlags(&sl.aux, cudaStreamNonBlocking) == cudaSuccess &&
             cudaEventCreateWithFlags(&sl.ev_fork, cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreateWithFlags(&sl.ev_join, cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreateWithFlags(&sl.ev_done, cudaEventDisableTiming) == cudaSuccess &&
             cudaHostAlloc((void**)&sl.hres, HRES_BYTES, cudaHostAllocMapped) == cudaSuccess &&
             cudaHostGetDevicePointer((void**)&sl.hres_dev, sl.hres, 0) == cudaSuccess;
    }
    if (!ok) {
        for (auto& sl : ctx->slots) {
            if (sl.stream) cudaStreamDestroy(sl.stream);
            if (sl.aux) cudaStreamDestroy(sl.aux);
            if (sl.ev_fork) cudaEventDestroy(sl.ev_fork);
            if (sl.ev_join) cudaEventDestroy(sl.ev_join);
            if (sl.ev_done) cudaEventDestroy(sl.ev_done);
            if (sl.hres) cudaFreeHost(sl.hres);
        }
        delete ctx;
        return BGLS_ERR_CUDA;
    }
    const char* eng = getenv("BGLS_ENGINE");
    ctx->thread_engine = eng && std::string(eng) == "thread";
    ctx->machine_only = eng && std::string(eng) == "machine";
    ctx->engine = ctx->thread_engine ? ENGINE_THREAD : ctx->machine_only ? ENGINE_MACHINE : (eng && std::string(eng) == "slot") ? ENGINE_SLOT : ENGINE_AUTO;
    const char* mil = getenv("BGLS_MILLER");
    ctx->miller16 = mil && std::string(mil) == "m16";
    if (const char* sp = getenv("BGLS_SLOT_SPLIT")) ctx->slot_split = atoi(sp) != 0 ? 1 : 0;
    if (const char* hm = getenv("BGLS_HASH")) ctx->hash_mode = std::string(hm) == "pool" ? 1 : std::string(hm) == "wide" ? 2 : 0;
    if (const char* ab = getenv("BGLS_AGG_BLOCKS")) ctx->agg_blocks_per_sm = atoi(ab);
    const char* mw = getenv("BGLS_MIN_WPB");
    if (mw && atoi(mw) > 0) ctx->min_wpb = atoi(mw);
    ctx->no_shared = mil && std::string(mil) == "noshare";
    cudaDeviceGetAttribute(&ctx->sms, cudaDevAttrMultiProcessorCount, device);
    if (ctx->sms <= 0) ctx->sms = 148;
    int rc = upload_mach<mtab::BN254_M, mtab::BN254_M_T, mtab::BN254_F, mtab::BN254_F_T, mtab::BN254_MP, mtab::BN254_MP_T>(ctx, ctx->mach[0]);
    if (!rc) rc = upload_mach<mtab::BLS381_M, mtab::BLS381_M_T, mtab::BLS381_F, mtab::BLS381_F_T, mtab::BLS381_MP, mtab::BLS381_MP_T>(ctx, ctx->mach[1]);
    if (getenv("BGLS_TRACE")) {
        if (cudaMalloc((void**)&ctx->trace, (3 * (1ull << 20) + 1) * 8) != cudaSuccess) ctx->trace = nullptr;
        else cudaMemset(ctx->trace, 0, 8);
    }
    if (!rc) rc = upload_slot<BN254, SlotBN, mtab::BN254_F>(ctx, ctx->sloteng[0]);
    if (!rc) rc = upload_slot<BLS381, SlotBLS, mtab::BLS381_F>(ctx, ctx->sloteng[1]);
    if (rc) {
        bgls_ctx_destroy(ctx);
        return rc;
    }
    *out = ctx;
    return BGLS_OK;
}
void bgls_ctx_destroy(bgls_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    for (auto& sl : ctx->slots) {
        if (sl.scratch) cudaFree(sl.scratch);
        if (sl.tickets) cudaFree(sl.tickets);
        if (sl.hres) cudaFreeHost(sl.hres);
        if (sl.stream) cudaStreamDestroy(sl.stream);
        if (sl.aux) cudaStreamDestroy(sl.aux);
        if (sl.ev_fork) cudaEventDestroy(sl.ev_fork);
        if (sl.ev_join) cudaEventDestroy(sl.ev_join);
        if (sl.ev_done) cudaEventDestroy(sl.ev_done);
    }
    for (auto& d : ctx->mach)
        if (d.blob) cudaFree(d.blob);
    for (auto& d : ctx->sloteng)
        if (d.blob) cudaFree(d.blob);
    if (ctx->trace) {
        const char* path = getenv("BGLS_TRACE");
        unsigned long long cnt = 0;
        cudaMemcpy(&cnt, ctx->trace, 8, cudaMemcpyDeviceToHost);
        if (cnt > (1ull << 20)) cnt = 1ull << 20;
        std::vector<unsigned long long> h(3 * cnt);
        if (cnt) cudaMemcpy(h.data(), ctx->trace + 1, 3 * cnt * 8, cudaMemcpyDeviceToHost);
        if (FILE* f = path ? fopen(path, "w") : nullptr) {
            for (unsigned long long i = 0; i < cnt; i++) fprintf(f, "%llu %llu %llu\n", h[3 * i], h[3 * i + 1], h[3 * i + 2]);
            fclose(f);
        }
        cudaFree(ctx->trace);
    }
    for (int r = 0; r < (int)ctx->xch.peer.size(); r++)
        if (ctx->xch.peer[r] && r != ctx->xch.rank) cudaIpcCloseMemHandle(ctx->xch.peer[r]);
    if (ctx->xch.local) cudaFree(ctx->xch.local);
    if (ctx->xch.d_err) cudaFree(ctx->xch.d_err);
    for (int i = 0; i < 3; i++)
        if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
    delete ctx;
}
const char* bgls_last_error(const bgls_ctx* ctx) {
    if (!ctx) return "null context";
    // a copy per calling thread, taken under the lock: the pointer stays valid (until this thread asks again) while other
    // threads fail and overwrite the context's text
    static thread_local std::string copy;
    {
        std::lock_guard<std::mutex> lk(const_cast<bgls_ctx*>(ctx)->err_mu);
        copy = ctx->err;
    }
    return copy.c_str();
}
uint64_t bgls_launch_count(const bgls_ctx* ctx) { return ctx ? ctx->launches.load() : 0; }

int bgls_set_profiling(bgls_ctx* ctx, int on) {
    if (!ctx) return BGLS_ERR_ARG;
    SlotLock sl(ctx);
    CU(cudaSetDevice(ctx->device));
    if (on && !ctx->ev[0])
        for (int i = 0; i < 3; i++) CU(cudaEventCreate(&ctx->ev[i]));
    ctx->profiling = on != 0;
    return BGLS_OK;
}
int bgls_last_kernel_ms(bgls_ctx* ctx, float* ms_main, float* ms_finish) {
    if (!ctx || !ctx->ev[0] || !ms_main || !ms_finish) return fail(ctx, BGLS_ERR_ARG, "profiling not enabled");
    SlotLock sl(ctx);
    CU(cudaSetDevice(ctx->device));
    CU(cudaEventSynchronize(ctx->ev[2]));
    CU(cudaEventElapsedTime(ms_main, ctx->ev[0], ctx->ev[1]));
    CU(cudaEventElapsedTime(ms_finish, ctx->ev[1], ctx->ev[2]));
    return BGLS_OK;
}
int bgls_intpipe_peak(bgls_ctx* ctx, double* wide_mac_per_s) {
    if (!ctx || !wide_mac_per_s) return fail(ctx, BGLS_ERR_ARG, "bad argument");
    SlotLock sl(ctx);
    CU(cudaSetDevice(ctx->device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, ctx->device));
    const int blocks = prop.multiProcessorCount * 2, threads = 1024, iters = 8192;
    int rc = ensure_scratch(ctx, sl.s, (size_t)blocks * threads * 4);
    if (rc) return rc;
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0));
    CU(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) {
        CU(cudaEventRecord(e0, sl.s->stream));
        k_intpipe_peak<<<blocks, threads, 0, sl.s->stream>>>((uint32_t*)sl.s->scratch, 3, iters);
        CU(cudaEventRecord(e1, sl.s->stream));
        CU(cudaEventSynchronize(e1));
        float ms;
        CU(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *wide_mac_per_s = (double)blocks * threads * iters * 8 / (best * 1e-3);
    return BGLS_OK;
}

// ---- host-buffer entry points
struct Busy {   // load estimate of use_slot_engine / hash_dev: host-buffer calls currently inside the library
    bgls_ctx* c;
    explicit Busy(bgls_ctx* c_) : c(c_) { c->host_calls++; }
    ~Busy() { c->host_calls--; }
};
static int pairing_host(bgls_ctx* ctx, int curve, const uint8_t* g1, const uint8_t* g2, size_t n, int do_final,
                        uint8_t* out, int* flag) {
    if (!ctx || !curve_ok(curve) || !out || (n && (!g1 || !g2))) return fail(ctx, BGLS_ERR_ARG, "bad argument");
    Busy busy(ctx);
    SlotLock sl(ctx);
    CU(cudaSetDevice(ctx->device));
    const size_t F = fp_bytes(curve), b1 = align_up(n * 2 * F), b2 = align_up(n * 4 * F), bo = align_up(12 * F + 16);
    const size_t bw = pairing_work_bytes(ctx, curve, n);
    int rc = ensure_scratch(ctx, sl.s, b1 + b2 + bo + bw);
    if (rc) return rc;
    uint8_t* base = (uint8_t*)sl.s->scratch;
    uint8_t *d1 = base, *d2 = base + b1, *dout = base + b1 + b2, *work = dout + bo;
    int* dflag = (int*)(dout + 12 * F);
    if (n) {
        CU(cudaMemcpyAsync(d1, g1, n * 2 * F, cudaMemcpyHostToDevice, sl.s->stream));
        CU(cudaMemcpyAsync(d2, g2, n * 4 * F, cudaMemcpyHostToDevice, sl.s->stream));
    }
    // the result (GT bytes + verdict, 388 / 580 bytes) is written by the last kernel straight into the slot's mapped pinned
    // buffer: no device-to-host copy is enqueued behind the kernels (a dependent entry that waits for milliseconds stalls
    // the hardware queue its stream shares with the streams of other calls)
    (void)dout; (void)dflag;
    uint8_t* hbuf = sl.s->hres;
    rc = pairing_dev(ctx, curve, d1, d2, n, do_final, sl.s->hres_dev, sl.s->hres_dev + 12 * F, work, sl.s->stream);
    if (rc) return rc;
    CU(cudaStreamSynchronize(sl.s->stream));
    memcpy(out, hbuf, 12 * F);
    if (flag) memcpy(flag, hbuf + 12 * F, 4);
    return BGLS_OK;
}
int bgls_pairing_product(bgls_ctx* ctx, int curve, const uint8_t* g1, const uint8_t* g2, size_t n, uint8_t* out_gt, int* is_identity) {
    return pairing_host(ctx, curve, g1, g2, n, 1, out_gt, is_identity);
}
int bgls_pair(bgls_ctx* ctx, int curve, const uint8_t* g1, const uint8_t* g2, uint8_t* out_gt) {
    return pairing_host(ctx, curve, g1, g2, 1, 1, out_gt, nullptr);
}
int bgls_miller_product(bgls_ctx* ctx, int curve, const uint8_t* g1, const uint8_t* g2, size_t n, uint8_t* out_f) {
    return pairing_host(ctx, curve, g1, g2, n, 0, out_f, nullptr);
}
int bgls_final_exp_product(bgls_ctx* ctx, int curve, const uint8_t* partials, size_t k, uint8_t* out_gt, int* is_identity) {
    if (!ctx || !curve_ok(curve) || !out_gt || (k && !partials)) return fail(ctx, BGLS_ERR_ARG, "bad argument");
    SlotLock sl(ctx);
    CU(cudaSetDevice(ctx->device));
    const size_t F = fp_bytes(curve), bi = align_up(k * 12 * F), bo = align_up(12 * F + 16);
    int rc = ensure_scratch(ctx, sl.s, bi + bo + mach_work_for(curve, k));
    if (rc) return rc;
    uint8_t *din = (uint8_t*)sl.s->scratch, *dout = din + bi;
    int* dflag = (int*)(dout + 12 * F);
    if (k) CU(cudaMemcpyAsync(din, partials, k * 12 * F, cudaMemcpyHostToDevice, sl.s->stream));
    rc = finish_bytes_dev(ctx, curve, din, k, 1, dout, dflag, dout + bo, sl.s->stream);
    if (rc) return rc;
    uint8_t* hbuf = sl.s->hres;
    CU(cudaMemcpyAsync(hbuf, dout, 12 * F + 4, cudaMemcpyDeviceToHost, sl.s->stream));
    CU(cudaStreamSynchronize(sl.s->stream));
    memcpy(out_gt, hbuf, 12 * F);
    if (is_identity) memcpy(is_identity, hbuf + 12 * F, 4);
    return BGLS_OK;
}
int bgls_gt_mul(bgls_ctx* ctx, int curve, const uint8_t* a, const uint8_t* b, uint8_t* out_gt) {
    if (!ctx || !curve_ok(curve) || !a || !b || !out_gt) return fail(ctx, BGLS_ERR_ARG, "bad argument");
    SlotLock sl(ctx);
    CU(cudaSetDevice(ctx->device));
    const size_t F = fp_bytes(curve), bi = align_up(2 * 12 * F), bo = align_up(12 * F + 16);
    int rc = ensure_scratch(ctx, sl.s, bi + bo + mach_work_for(curve, 2));
    if (rc) return rc;
    uint8_t *din = (uint8_t*)sl.s->scratch, *dout = din + bi;
    CU(cudaMemcpyAsync(din, a, 12 * F, cudaMemcpyHostToDevice, sl.s->stream));
    CU(cudaMemcpyAsync(din + 12 * F, b, 12 * F, cudaMemcpyHostToDevice, sl.s->stream));
    rc = finish_bytes_dev(ctx, curve, din, 2, 0, dout, nullptr, dout + bo, sl.s->stream);
    if (rc) return rc;
    CU(cudaMemcpyAsync(sl.s->hres, dout, 12 * F, cudaMemcpyDeviceToHost, sl.s->stream));
    CU(cudaStreamSynchronize(sl.s->stream));
    memcpy(out_gt, sl.s->hres, 12 * F);
    return BGLS_OK;
}
int bgls_aggregate_points(bgls_ctx* ctx, int curve, int group, const uint8_t* pts, size_t n, uint8_t* out) {
    if (!ctx || !curve_ok(curve) || (group != 1 && group != 2) || !pts || !out || n == 0) return fail(ctx, BGLS_ERR_ARG, "bad argument");
    SlotLock sl(ctx);
    CU(cudaSetDevice(ctx->device));
    const size_t rec = 2 * group * fp_bytes(curve), bi = align_up(n * rec), bo = align_up(rec);
    const size_t bw = agg_work_bytes(curve, group, n);
    int rc = ensure_scratch(ctx, sl.s, bi + bo + bw);
    if (rc) return rc;
    uint8_t *din = (uint8_t*)sl.s->scratch, *dout = din + bi, *work = dout + bo;
    CU(cudaMemcpyAsync(din, pts, n * rec, cudaMemcpyHostToDevice, sl.s->stream));
    rc = aggregate_dev(ctx, curve, group, din, n, dout, work, sl.s->stream);
    if (rc) return rc;
    CU(cudaMemcpyAsync(sl.s->hres, dout, rec, cudaMemcpyDeviceToHost, sl.s->stream));
    CU(cudaStreamSynchronize(sl.s->stream));
    memcpy(out, sl.s->hres, rec);
    return BGLS_OK;
}
int bgls_scale_points(bgls_ctx* ctx, int curve, int group, const uint8_t* pts, const uint8_t* scalars, size_t n, uint8_t* out) {
    if (!ctx || !curve_ok(curve) || (group != 1 && group != 2) || (n && (!pts || !scalars || !out))) return fail(ctx, BGLS_ERR_ARG, "bad argument");
    if (n == 0) return BGLS_OK;
    SlotLock sl(ctx);
    CU(cudaSetDevice(ctx->device));
    const size_t rec = 2 * group * fp_bytes(curve), bi = align_up(n * rec), bs = align_up(n * 32);
    int rc = ensure_scratch(ctx, sl.s, 2 * bi + bs);
    if (rc) return rc;
    uint8_t *din = (uint8_t*)sl.s->scratch, *dsc = din + bi, *dout = dsc + bs;
    CU(cudaMemcpyAsync(din, pts, n * rec, cudaMemcpyHostToDevice, sl.s->stream));
    CU(cudaMemcpyAsync(dsc, scalars, n * 32, cudaMemcpyHostToDevice, sl.s->stream));
    rc = scale_dev(ctx, curve, group, din, dsc, n, dout, sl.s->stream);
    if (rc) return rc;
    CU(cudaMemcpyAsync(out, dout, n * rec, cudaMemcpyDeviceToHost, sl.s->stream));
    CU(cudaStreamSynchronize(sl.s->stream));
    return BGLS_OK;
}
static int codec_dev(bgls_ctx* ctx, int curve, int group, bool compress, const void* d_in, size_t n, int check_subgroup,
                     void* d_out, void* d_ok, cudaStream_t s) {
    if (n == 0) return BGLS_OK;
    const unsigned nb = (unsigned)((n + TB - 1) / TB);
    const uint8_t* in = (const uint8_t*)d_in;
    uint8_t *out = (uint8_t*)d_out, *ok = (uint8_t*)d_ok;
    if (compress) {
        if (curve == BGLS_ALTBN128) {
            if (group == 1) k_compress<BN254, 1><<<nb, TB, 0, s>>>(in, n, out);
            else k_compress<BN254, 2><<<nb, TB, 0, s>>>(in, n, out);
        } else {
            if (group == 1) k_compress<BLS381, 1><<<nb, TB, 0, s>>>(in, n, out);
            else k_compress<BLS381, 2><<<nb, TB, 0, s>>>(in, n, out);
        }
    } else {
        if (curve == BGLS_ALTBN128) {
            if (group == 1) k_decompress<BN254, 1><<<nb, TB, 0, s>>>(in, n, check_subgroup, out, ok);
            else k_decompress<BN254, 2><<<nb, TB, 0, s>>>(in, n, check_subgroup, out, ok);
        } else {
            if (group == 1) k_decompress<BLS381, 1><<<nb, TB, 0, s>>>(in, n, check_subgroup, out, ok);
            else k_decompress<BLS381, 2><<<nb, TB, 0, s>>>(in, n, check_subgroup, out, ok);
        }
    }
    ctx->launches++;
    CU(cudaGetLastError());
    return BGLS_OK;
}
int bgls_compress_points(bgls_ctx* ctx, int curve, int group, const uint8_t* pts, size_t n, uint8_t* out) {
    if (!ctx || !curve_ok(curve) || (group != 1 && group != 2) || (n && (!pts || !out))) return fail(ctx, BGLS_ERR_ARG, "bad argument");
    if (n == 0) return BGLS_OK;
    SlotLock sl(ctx);
    CU(cudaSetDevice(ctx->device));
    const size_t F = fp_bytes(curve), bi = align_up(n * 2 * group * F), bo = align_up(n * group * F);
    int rc = ensure_scratch(ctx, sl.s, bi + bo);
    if (rc) return rc;
    uint8_t *din = (uint8_t*)sl.s->scratch, *dout = din + bi;
    CU(cudaMemcpyAsync(din, pts, n * 2 * group * F, cudaMemcpyHostToDevice, sl.s->stream));
    rc = codec_dev(ctx, curve, group, true, din, n, 0, dout, nullptr, sl.s->stream);
    if (rc) return rc;
    CU(cudaMemcpyAsync(out, dout, n * group * F, cudaMemcpyDeviceToHost, sl.s->stream));
    CU(cudaStreamSynchronize(sl.s->stream));
    return BGLS_OK;
}
int bgls_decompress_points(bgls_ctx* ctx, int curve, int group, const uint8_t* in, size_t n, int check_subgroup,
                           uint8_t* out_pts, uint8_t* out_ok) {
    if (!ctx || !curve_ok(curve) || (group != 1 && group != 2) || (n && (!in || !out_pts || !out_ok))) return fail(ctx, BGLS_ERR_ARG, "bad argument");
    if (n == 0) return BGLS_OK;
    SlotLock sl(ctx);
    CU(cudaSetDevice(ctx->device));
    const size_t F = fp_bytes(curve), bi = align_up(n * group * F), bo = align_up(n * 2 * group * F), bk = align_up(n);
    int rc = ensure_scratch(ctx, sl.s, bi + bo + bk);
    if (rc) return rc;
    uint8_t *din = (uint8_t*)sl.s->scratch, *dout = din + bi, *dok = dout + bo;
    CU(cudaMemcpyAsync(din, in, n * group * F, cudaMemcpyHostToDevice, sl.s->stream));
    rc = codec_dev(ctx, curve, group, false, din, n, check_subgroup, dout, dok, sl.s->stream);
    if (rc) return rc;
    CU(cudaMemcpyAsync(out_pts, dout, n * 2 * group * F, cudaMemcpyDeviceToHost, sl.s->stream));
    CU(cudaMemcpyAsync(out_ok, dok, n, cudaMemcpyDeviceToHost, sl.s->stream));
    CU(cudaStreamSynchronize(sl.s->stream));
    return BGLS_OK;
}
int bgls_compress_points_dev(bgls_ctx* ctx, int curve, int group, const void* d_pts, size_t n, void* d_out, void* stream) {
    if (!ctx || !curve_ok(curve) || (group != 1 && group != 2) || (n && (!d_pts || !d_out))) return fail(ctx, BGLS_ERR_ARG, "bad argument");
    CU(cudaSetDevice(ctx->device));
    return codec_dev(ctx, curve, group, true, d_pts, n, 0, d_out, nullptr, (cudaStream_t)stream);
}
int bgls_decompress_points_dev(bgls_ctx* ctx, int curve, int group, const void* d_in, size_t n, int check_subgroup,
                               void* d_out_pts, void* d_out_ok, void* stream) {
    if (!ctx || !curve_ok(curve) || (group != 1 && group != 2) || (n && (!d_in || !d_out_pts || !d_out_ok))) return fail(ctx, BGLS_ERR_ARG, "bad argument");
    CU(cudaSetDevice(ctx->device));
    return codec_dev(ctx, curve, group, false, d_in, n, check_subgroup, d_out_pts, d_out_ok, (cudaStream_t)stream);
}
static int validate_dev(bgls_ctx* ctx, int curve, int group, const void* d_pts, size_t n, int mode, void* d_ok, cudaStream_t s) {
    if (n == 0) return BGLS_OK;
    const unsigned nb = (unsigned)((n + TB - 1) / TB);
    const uint8_t* in = (const uint8_t*)d_pts;
    uint8_t* ok = (uint8_t*)d_ok;
    const int sub = mode == BGLS_VALIDATE_REFERENCE ? 1 : 0;
    if (curve == BGLS_ALTBN128) {
        if (group == 1) k_validate<BN254, 1><<<nb, TB, 0, s>>>(in, n, sub, ok);
        else k_validate<BN254, 2><<<nb, TB, 0, s>>>(in, n, sub, ok);
    } else {
        if (group == 1) k_validate<BLS381, 1><<<nb, TB, 0, s>>>(in, n, sub, ok);
        else k_validate<BLS381, 2><<<nb, TB, 0, s>>>(in, n, sub, ok);
    }
    ctx->launches++;
    CU(cudaGetLastError());
    return BGLS_OK;
}
int bgls_validate_points(bgls_ctx* ctx, int curve, int group, const uint8_t* pts, size_t n, int mode, uint8_t* out_ok) {
    if (!ctx || !curve_ok(curve) || (group != 1 && group != 2) || (mode != BGLS_VALIDATE_ONCURVE && mode != BGLS_VALIDATE_REFERENCE) ||
        (n && (!pts || !out_ok)))
        return fail(ctx, BGLS_ERR_ARG, "bad argument");
    if (n == 0) return BGLS_OK;
    SlotLock sl(ctx);
    CU(cudaSetDevice(ctx->device));
    const size_t rec = 2 * group * fp_bytes(curve), bi = align_up(n * rec), bk = align_up(n);
    int rc = ensure_scratch(ctx, sl.s, bi + bk);
    if (rc) return rc;
    uint8_t *din = (uint8_t*)sl.s->scratch, *dok = din + bi;
    CU(cudaMemcpyAsync(din, pts, n * rec, cudaMemcpyHostToDevice, sl.s->stream));
    rc = validate_dev(ctx, curve, group, din, n, mode, dok, sl.s->stream);
    if (rc) return rc;
    CU(cudaMemcpyAsync(out_ok, dok, n, cudaMemcpyDeviceToHost, sl.s->stream));
    CU(cudaStreamSynchronize(sl.s->stream));
    return BGLS_OK;
}
int bgls_validate_points_dev(bgls_ctx* ctx, int curve, int group, const void* d_pts, size_t n, int mode, void* d_out_ok, void* stream) {
    if (!ctx || !curve_ok(curve) || (group != 1 && group != 2) || (mode != BGLS_VALIDATE_ONCURVE && mode != BGLS_VALIDATE_REFERENCE) ||
        (n && (!d_pts || !d_out_ok)))
        return fail(ctx, BGLS_ERR_ARG, "bad argument");
    CU(cudaSetDevice(ctx->device));
    return validate_dev(ctx, curve, group, d_pts, n, mode, d_out_ok, (cudaStream_t)stream);
}
int bgls_gt_pow(bgls_ctx* ctx, int curve, const uint8_t* a, const uint8_t* exponent32, int negative, uint8_t* out_gt) {
    if (!ctx || !curve_ok(curve) || !a || !exponent32 || !out_gt) return fail(ctx, BGLS_ERR_ARG, "bad argument");
    SlotLock sl(ctx);
    CU(cudaSetDevice(ctx->device));
    const size_t F = fp_bytes(curve), bi = align_up(12 * F), be = align_up(32);
    int rc = ensure_scratch(ctx, sl.s, 2 * bi + be);
    if (rc) return rc;
    uint8_t *da = (uint8_t*)sl.s->scratch, *de = da + bi, *dout = de + be;
    CU(cudaMemcpyAsync(da, a, 12 * F, cudaMemcpyHostToDevice, sl.s->stream));
    CU(cudaMemcpyAsync(de, exponent32, 32, cudaMemcpyHostToDevice, sl.s->stream));
    if (curve == BGLS_ALTBN128) k_gt_pow<BN254><<<1, 32, 0, sl.s->stream>>>(da, de, negative, dout);
    else k_gt_pow<BLS381><<<1, 32, 0, sl.s->stream>>>(da, de, negative, dout);
    ctx->launches++;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(sl.s->hres, dout, 12 * F, cudaMemcpyDeviceToHost, sl.s->stream));
    CU(cudaStreamSynchronize(sl.s->stream));
    memcpy(out_gt, sl.s->hres, 12 * F);
    return BGLS_OK;
}
static int hash_dev(bgls_ctx* ctx, int curve, const void* d_msgs, const void* d_off, size_t n, void* d_out, cudaStream_t s) {
    if (n == 0) return BGLS_OK;
    if (curve == BGLS_ALTBN128) {
        // several verifications in flight (or one large batch of messages): the pooled form does a third of the work;
        // a lone verification keeps the one-round latency form.  BGLS_HASH=pool|wide overrides.
        const bool pool = ctx->hash_mode == 1 || (ctx->hash_mode == 0 && n >= 64 && (busy_estimate(ctx) >= 3 || n >= 16384));
        if (pool)
            k_hash_to_g1_bn_pool<BN254><<<(unsigned)((n + 31) / 32), TB, 0, s>>>((const uint8_t*)d_msgs, (const unsigned long long*)d_off, n, (uint8_t*)d_out);
        else
            k_hash_to_g1_bn<BN254><<<(unsigned)((n * 8 + TB - 1) / TB), TB, 0, s>>>((const uint8_t*)d_msgs, (const unsigned long long*)d_off, n, (uint8_t*)d_out);
    } else {
        const bool pool = ctx->hash_mode == 1 || (ctx->hash_mode == 0 && n >= 64 && (busy_estimate(ctx) >= 3 || n >= 16384));
        if (pool)
            k_hash_to_g1_bls_one<BLS381><<<(unsigned)((n + TB - 1) / TB), TB, 0, s>>>((const uint8_t*)d_msgs, (const unsigned long long*)d_off, n, (uint8_t*)d_out);
        else
            k_hash_to_g1_bls<BLS381><<<(unsigned)((n * 2 + TB - 1) / TB), TB, 0, s>>>((const uint8_t*)d_msgs, (const unsigned long long*)d_off, n, (uint8_t*)d_out);
    }
    ctx->launches++;
    CU(cudaGetLastError());
    return BGLS_OK;
}
int bgls_hash_to_g1(bgls_ctx* ctx, int curve, const uint8_t* msgs, const uint64_t* offsets, size_t n, uint8_t* out) {
    if (!ctx || !curve_ok(curve) || !offsets || (n && !out)) return fail(ctx, BGLS_ERR_ARG, "bad argument");
    if (n == 0) return BGLS_OK;
    for (size_t i = 0; i < n; i++)
        if (offsets[i + 1] < offsets[i]) return fail(ctx, BGLS_ERR_ARG, "offsets not monotone");
    const size_t total = offsets[n] - offsets[0];
    if (total && !msgs) return fail(ctx, BGLS_ERR_ARG, "bad argument");
    SlotLock sl(ctx);
    CU(cudaSetDevice(ctx->device));
    const size_t F = fp_bytes(curve), bm = align_up(total + 16), bo = align_up((n + 1) * 8), bp = align_up(n * 2 * F);
    int rc = ensure_scratch(ctx, sl.s, bm + bo + bp);
    if (rc) return rc;
    uint8_t *dm = (uint8_t*)sl.s->scratch, *doff = dm + bm, *dout = doff + bo;
    std::vector<uint64_t> rel(n + 1);
    for (size_t i = 0; i <= n; i++) rel[i] = offsets[i] - offsets[0];
    if (total) CU(cudaMemcpyAsync(dm, msgs + offsets[0], total, cudaMemcpyHostToDevice, sl.s->stream));
    CU(cudaMemcpyAsync(doff, rel.data(), (n + 1) * 8, cudaMemcpyHostToDevice, sl.s->stream));
    rc = hash_dev(ctx, curve, dm, doff, n, dout, sl.s->stream);
    if (rc) return rc;
    CU(cudaMemcpyAsync(out, dout, n * 2 * F, cudaMemcpyDeviceToHost, sl.s->stream));
    CU(cudaStreamSynchronize(sl.s->stream));
    return BGLS_OK;
}
int bgls_hash_to_g1_dev(bgls_ctx* ctx, int curve, const void* d_msgs, const void* d_offsets, size_t n, void* d_out, void* stream) {
    if (!ctx || !curve_ok(curve) || (n && (!d_offsets || !d_out))) return fail(ctx, BGLS_ERR_ARG, "bad argument");
    CU(cudaSetDevice(ctx->device));
    return hash_dev(ctx, curve, d_msgs, d_offsets, n, d_out, (cudaStream_t)stream);
}
// ---- verifyAggSig as one call (bgls/bgls.go:94-119)
// G2 generators (curves/altbn128_test.go:26-35, curves/bls12_381.go:328-346) and field primes, wire form (big-endian)
static const uint8_t kG2GenBN254[128] = {0x19, 0x8e, 0x93, 0x93, 0x92, 0x0d, 0x48, 0x3a, 0x72, 0x60, 0xbf, 0xb7, 0x31, 0xfb, 0x5d, 0x25, 0xf1, 0xaa, 0x49, 0x33, 0x35, 0xa9, 0xe7, 0x12, 0x97, 0xe4, 0x85, 0xb7, 0xae, 0xf3, 0x12, 0xc2, 0x18, 0x00, 0xde, 0xef, 0x12, 0x1f, 0x1e, 0x76, 0x42, 0x6a, 0x00, 0x66, 0x5e, 0x5c, 0x44, 0x79, 0x67, 0x43, 0x22, 0xd4, 0xf7, 0x5e, 0xda, 0xdd, 0x46, 0xde, 0xbd, 0x5c, 0xd9, 0x92, 0xf6, 0xed, 0x09, 0x06, 0x89, 0xd0, 0x58, 0x5f, 0xf0, 0x75, 0xec, 0x9e, 0x99, 0xad, 0x69, 0x0c, 0x33, 0x95, 0xbc, 0x4b, 0x31, 0x33, 0x70, 0xb3, 0x8e, 0xf3, 0x55, 0xac, 0xda, 0xdc, 0xd1, 0x22, 0x97, 0x5b, 0x12, 0xc8, 0x5e, 0xa5, 0xdb, 0x8c, 0x6d, 0xeb, 0x4a, 0xab, 0x71, 0x80, 0x8d, 0xcb, 0x40, 0x8f, 0xe3, 0xd1, 0xe7, 0x69, 0x0c, 0x43, 0xd3, 0x7b, 0x4c, 0xe6, 0xcc, 0x01, 0x66, 0xfa, 0x7d, 0xaa};
static const uint8_t kG2GenBLS381[192] = {0x13, 0xe0, 0x2b, 0x60, 0x52, 0x71, 0x9f, 0x60, 0x7d, 0xac, 0xd3, 0xa0, 0x88, 0x27, 0x4f, 0x65, 0x59, 0x6b, 0xd0, 0xd0, 0x99, 0x20, 0xb6, 0x1a, 0xb5, 0xda, 0x61, 0xbb, 0xdc, 0x7f, 0x50, 0x49, 0x33, 0x4c, 0xf1, 0x12, 0x13, 0x94, 0x5d, 0x57, 0xe5, 0xac, 0x7d, 0x05, 0x5d, 0x04, 0x2b, 0x7e, 0x02, 0x4a, 0xa2, 0xb2, 0xf0, 0x8f, 0x0a, 0x91, 0x26, 0x08, 0x05, 0x27, 0x2d, 0xc5, 0x10, 0x51, 0xc6, 0xe4, 0x7a, 0xd4, 0xfa, 0x40, 0x3b, 0x02, 0xb4, 0x51, 0x0b, 0x64, 0x7a, 0xe3, 0xd1, 0x77, 0x0b, 0xac, 0x03, 0x26, 0xa8, 0x05, 0xbb, 0xef, 0xd4, 0x80, 0x56, 0xc8, 0xc1, 0x21, 0xbd, 0xb8, 0x06, 0x06, 0xc4, 0xa0, 0x2e, 0xa7, 0x34, 0xcc, 0x32, 0xac, 0xd2, 0xb0, 0x2b, 0xc2, 0x8b, 0x99, 0xcb, 0x3e, 0x28, 0x7e, 0x85, 0xa7, 0x63, 0xaf, 0x26, 0x74, 0x92, 0xab, 0x57, 0x2e, 0x99, 0xab, 0x3f, 0x37, 0x0d, 0x27, 0x5c, 0xec, 0x1d, 0xa1, 0xaa, 0xa9, 0x07, 0x5f, 0xf0, 0x5f, 0x79, 0xbe, 0x0c, 0xe5, 0xd5, 0x27, 0x72, 0x7d, 0x6e, 0x11, 0x8c, 0xc9, 0xcd, 0xc6, 0xda, 0x2e, 0x35, 0x1a, 0xad, 0xfd, 0x9b, 0xaa, 0x8c, 0xbd, 0xd3, 0xa7, 0x6d, 0x42, 0x9a, 0x69, 0x51, 0x60, 0xd1, 0x2c, 0x92, 0x3a, 0xc9, 0xcc, 0x3b, 0xac, 0xa2, 0x89, 0xe1, 0x93, 0x54, 0x86, 0x08, 0xb8, 0x28, 0x01};
static const uint8_t kPrimeBN254[32] = {0x30, 0x64, 0x4e, 0x72, 0xe1, 0x31, 0xa0, 0x29, 0xb8, 0x50, 0x45, 0xb6, 0x81, 0x81, 0x58, 0x5d, 0x97, 0x81, 0x6a, 0x91, 0x68, 0x71, 0xca, 0x8d, 0x3c, 0x20, 0x8c, 0x16, 0xd8, 0x7c, 0xfd, 0x47};
static const uint8_t kPrimeBLS381[48] = {0x1a, 0x01, 0x11, 0xea, 0x39, 0x7f, 0xe6, 0x9a, 0x4b, 0x1b, 0xa7, 0xb6, 0x43, 0x4b, 0xac, 0xd7, 0x64, 0x77, 0x4b, 0x84, 0xf3, 0x85, 0x12, 0xbf, 0x67, 0x30, 0xd2, 0xa0, 0xf6, 0xb0, 0xf6, 0x24, 0x1e, 0xab, 0xff, 0xfe, 0xb1, 0x53, 0xff, 0xff, 0xb9, 0xfe, 0xff, 0xff, 0xff, 0xff, 0xaa, 0xab};

// y -> q - y on a big-endian field element (Point.Mul(-1) of the aggregate signature, bgls/bgls.go:112); 0 stays 0
static void be_negate(uint8_t* y, const uint8_t* q, size_t F) {
    bool zero = true;
    for (size_t i = 0; i < F; i++) zero = zero && y[i] == 0;
    if (zero) return;
    int borrow = 0;
    for (size_t i = F; i-- > 0;) {
        int d = (int)q[i] - (int)y[i] - borrow;
        borrow = d < 0;
        y[i] = (uint8_t)(d + (borrow ? 256 : 0));
    }
}
int bgls_verify_aggregate_signature(bgls_ctx* ctx, int curve, const uint8_t* msgs, const uint64_t* offsets, size_t n,
                                    const uint8_t* keys, const uint8_t* sig, int allow_duplicates, int* ok) {
    if (!ctx || !curve_ok(curve) || !offsets || !sig || !ok || (n && !keys)) return fail(ctx, BGLS_ERR_ARG, "bad argument");
    for (size_t i = 0; i < n; i++)
        if (offsets[i + 1] < offsets[i]) return fail(ctx, BGLS_ERR_ARG, "offsets not monotone");
    const size_t total = offsets[n] - offsets[0];
    if (total && !msgs) return fail(ctx, BGLS_ERR_ARG, "bad argument");
    *ok = 0;
    if (!allow_duplicates) {   // containsDuplicateMessage, bgls/bgls.go:98-102,139-150: the verdict is false before any pairing
        std::unordered_set<std::string> seen;
        seen.reserve(n * 2);
        const char* base = msgs ? (const char*)msgs : "";
        for (size_t i = 0; i < n; i++)
            if (!seen.emplace(base + (total ? offsets[i] : 0), offsets[i + 1] - offsets[i]).second) return BGLS_OK;
    }
    Busy busy(ctx);
    SlotLock sl(ctx);
    CU(cudaSetDevice(ctx->device));
    const size_t F = fp_bytes(curve), np = n + 1;
    const size_t bm = align_up(total + 16), bo = align_up((n + 1) * 8), b1 = align_up(np * 2 * F), b2 = align_up(np * 4 * F);
    const size_t bout = align_up(12 * F + 16), bw = pairing_work_bytes(ctx, curve, np);
    int rc = ensure_scratch(ctx, sl.s, bm + bo + b1 + b2 + bout + bw);
    if (rc) return rc;
    uint8_t* dm = (uint8_t*)sl.s->scratch;
    uint8_t *doff = dm + bm, *d1 = doff + bo, *d2 = d1 + b1, *dout = d2 + b2, *work = dout + bout;
    int* dflag = (int*)(dout + 12 * F);
    cudaStream_t s = sl.s->stream;
    std::vector<uint64_t> rel(n + 1);
    for (size_t i = 0; i <= n; i++) rel[i] = offsets[i] - offsets[0];
    uint8_t tail[2 * 48];   // -sigma
    memcpy(tail, sig, 2 * F);
    const bool sig_inf = (curve == BGLS_BLS12_381) && (tail[0] & 0x40);
    if (!sig_inf) be_negate(tail + F, curve == BGLS_ALTBN128 ? kPrimeBN254 : kPrimeBLS381, F);
    if (total) CU(cudaMemcpyAsync(dm, msgs + offsets[0], total, cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(doff, rel.data(), (n + 1) * 8, cudaMemcpyHostToDevice, s));
    if (n) CU(cudaMemcpyAsync(d2, keys, n * 4 * F, cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(d2 + n * 4 * F, curve == BGLS_ALTBN128 ? kG2GenBN254 : kG2GenBLS381, 4 * F, cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(d1 + n * 2 * F, tail, 2 * F, cudaMemcpyHostToDevice, s));
    rc = hash_dev(ctx, curve, dm, doff, n, d1, s);                       // pts1[i] = HashToG1(msgs[i]), bgls.go:106-111
    if (rc) return rc;
    (void)dout; (void)dflag;
    rc = pairing_dev(ctx, curve, d1, d2, np, 1, sl.s->hres_dev + 16, sl.s->hres_dev, work, s);    // PairingProduct(pts1, pts2), bgls.go:114
    if (rc) return rc;
    CU(cudaStreamSynchronize(s));   // also keeps `rel` and `tail` alive until the copies are done
    memcpy(ok, sl.s->hres, 4);      // aggPt.Equals(GetGTIdentity()), bgls.go:115-118
    return BGLS_OK;
}
int bgls_verify_multi_signature(bgls_ctx* ctx, int curve, const uint8_t* msg, size_t msg_len, const uint8_t* keys, size_t n,
                                const uint8_t* sig, int* ok) {
    if (!ctx || !curve_ok(curve) || !keys || !sig || !ok || n == 0 || (msg_len && !msg)) return fail(ctx, BGLS_ERR_ARG, "bad argument");
    *ok = 0;
    Busy busy(ctx);
    SlotLock sl(ctx);
    CU(cudaSetDevice(ctx->device));
    const size_t F = fp_bytes(curve);
    const size_t bm = align_up(msg_len + 16), bo = align_up(16), bk = align_up(n * 4 * F), b1 = align_up(2 * 2 * F), b2 = align_up(2 * 4 * F);
    const size_t bout = align_up(12 * F + 16);
    const size_t bw = std::max(agg_work_bytes(curve, 2, n), pairing_work_bytes(ctx, curve, 2));
    int rc = ensure_scratch(ctx, sl.s, bm + bo + bk + b1 + b2 + bout + bw);
    if (rc) return rc;
    uint8_t* dm = (uint8_t*)sl.s->scratch;
    uint8_t *doff = dm + bm, *dk = doff + bo, *d1 = dk + bk, *d2 = d1 + b1, *dout = d2 + b2, *work = dout + bout;
    int* dflag = (int*)(dout + 12 * F);
    cudaStream_t s = sl.s->stream;
    const uint64_t rel[2] = {0, (uint64_t)msg_len};
    uint8_t tail[2 * 48];
    memcpy(tail, sig, 2 * F);
    // the reference checks e(-H(m), vs) e(sigma, g2) == 1 (bgls.go:65-70); e(H(m), vs) e(-sigma, g2) is its inverse, so
    // the verdict is the same and the negation is one field subtraction on the host
    if (!((curve == BGLS_BLS12_381) && (tail[0] & 0x40))) be_negate(tail + F, curve == BGLS_ALTBN128 ? kPrimeBN254 : kPrimeBLS381, F);
    if (msg_len) CU(cudaMemcpyAsync(dm, msg, msg_len, cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(doff, rel, 16, cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(dk, keys, n * 4 * F, cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(d1 + 2 * F, tail, 2 * F, cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(d2 + 4 * F, curve == BGLS_ALTBN128 ? kG2GenBN254 : kG2GenBLS381, 4 * F, cudaMemcpyHostToDevice, s));
    // the hash of the message (one long dependent chain) and the key aggregation are independent: fork the hash onto
    // the slot's second stream and join before the pairing
    CU(cudaEventRecord(sl.s->ev_fork, s));
    CU(cudaStreamWaitEvent(sl.s->aux, sl.s->ev_fork, 0));
    rc = hash_dev(ctx, curve, dm, doff, 1, d1, sl.s->aux);         // HashToG1(msg), bgls.go:67
    if (rc) return rc;
    CU(cudaEventRecord(sl.s->ev_join, sl.s->aux));
    rc = aggregate_dev(ctx, curve, 2, dk, n, d2, work, s);        // vs = AggregatePoints(keys), bgls.go:90
    if (rc) return rc;
    CU(cudaStreamWaitEvent(s, sl.s->ev_join, 0));
    rc = pairing_dev(ctx, curve, d1, d2, 2, 1, dout, dflag, work, s);
    if (rc) return rc;
    CU(cudaMemcpyAsync(sl.s->hres, dflag, 4, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    memcpy(ok, sl.s->hres, 4);
    return BGLS_OK;
}
int bgls_pairing_check_batch(bgls_ctx* ctx, int curve, const uint8_t* g1, const uint8_t* g2, const uint64_t* offsets,
                             size_t nbatch, uint8_t* out_ok) {
    if (!ctx || !curve_ok(curve) || !offsets || (nbatch && !out_ok)) return fail(ctx, BGLS_ERR_ARG, "bad argument");
    if (nbatch == 0) return BGLS_OK;
    for (size_t b = 0; b < nbatch; b++)
        if (offsets[b + 1] < offsets[b]) return fail(ctx, BGLS_ERR_ARG, "offsets not monotone");
    if (offsets[0] != 0) return fail(ctx, BGLS_ERR_ARG, "offsets[0] must be 0");
    const size_t total = offsets[nbatch];
    if (total && (!g1 || !g2)) return fail(ctx, BGLS_ERR_ARG, "bad argument");
    SlotLock sl(ctx);
    CU(cudaSetDevice(ctx->device));
    const size_t F = fp_bytes(curve), b1 = align_up(total * 2 * F), b2 = align_up(total * 4 * F);
    const size_t bf = align_up((nbatch + 1) * 8), bk = align_up(nbatch);
    const size_t bw = std::max({align_up(total * fp12_dev_bytes(curve)), mach_work_for(curve, total + nbatch), slot_batch_work_bytes(curve, nbatch, total)});
    int rc = ensure_scratch(ctx, sl.s, b1 + b2 + bf + bk + bw);
    if (rc) return rc;
    uint8_t *d1 = (uint8_t*)sl.s->scratch, *d2 = d1 + b1, *doff = d2 + b2, *dok = doff + bf, *work = dok + bk;
    if (total) {
        CU(cudaMemcpyAsync(d1, g1, total * 2 * F, cudaMemcpyHostToDevice, sl.s->stream));
        CU(cudaMemcpyAsync(d2, g2, total * 4 * F, cudaMemcpyHostToDevice, sl.s->stream));
    }
    CU(cudaMemcpyAsync(doff, offsets, (nbatch + 1) * 8, cudaMemcpyHostToDevice, sl.s->stream));
    rc = batch_dev(ctx, curve, d1, d2, doff, nbatch, total, dok, work, sl.s->stream);
    if (rc) return rc;
    CU(cudaMemcpyAsync(out_ok, dok, nbatch, cudaMemcpyDeviceToHost, sl.s->stream));
    CU(cudaStreamSynchronize(sl.s->stream));
    return BGLS_OK;
}

// ---- device-resident entry points (no synchronisation)
static int dev_work(bgls_ctx* ctx, Slot* sl, size_t bytes, void** work) {
    int rc = ensure_scratch(ctx, sl, bytes);
    if (rc) return rc;
    *work = sl->scratch;
    return BGLS_OK;
}
int bgls_pairing_product_dev(bgls_ctx* ctx, int curve, const void* d_g1, const void* d_g2, size_t n, void* d_out_gt,
                             void* d_is_identity, void* stream) {
    if (!ctx || !curve_ok(curve) || !d_out_gt || (n && (!d_g1 || !d_g2))) return fail(ctx, BGLS_ERR_ARG, "bad argument");
    SlotLock sl(ctx, (cudaStream_t)stream);
    CU(cudaSetDevice(ctx->device));
    void* work;
    int rc = dev_work(ctx, sl.s, pairing_work_bytes(ctx, curve, n), &work);
    if (rc) return rc;
    return pairing_dev(ctx, curve, d_g1, d_g2, n, 1, d_out_gt, d_is_identity, work, (cudaStream_t)stream);
}
int bgls_miller_product_dev(bgls_ctx* ctx, int curve, const void* d_g1, const void* d_g2, size_t n, void* d_out_f, void* stream) {
    if (!ctx || !curve_ok(curve) || !d_out_f || (n && (!d_g1 || !d_g2))) return fail(ctx, BGLS_ERR_ARG, "bad argument");
    SlotLock sl(ctx, (cudaStream_t)stream);
    CU(cudaSetDevice(ctx->device));
    void* work;
    int rc = dev_work(ctx, sl.s, pairing_work_bytes(ctx, curve, n), &work);
    if (rc) return rc;
    return pairing_dev(ctx, curve, d_g1, d_g2, n, 0, d_out_f, nullptr, work, (cudaStream_t)stream);
}
// ---- peer-memory exchange (multi-GPU, one process per GPU): see k_exchange_send / k_exchange_wait
struct PeerPtrs {
    uint8_t* slot[8];
    unsigned long long* flag[8];
};
// records of one (lane, parity) are contiguous with the curve's own record size (the finishing side reads them as one
// array of `world` wire records); every (lane, parity) group starts on a multiple of XREC * world bytes
static size_t xch_rec_off(const Exchange& x, int lane, int parity, int r, size_t rec) {
    return (size_t)(lane * 2 + parity) * x.world * XREC + (size_t)r * rec;
}
static size_t xch_flag_base(const Exchange& x) { return align_up((size_t)x.lanes * 2 * x.world * XREC, 256); }
static size_t xch_flag_off(const Exchange& x, int lane, int parity, int r) {
    return xch_flag_base(x) + ((size_t)(lane * 2 + parity) * x.world + r) * sizeof(unsigned long long);
}
int bgls_exchange_create(bgls_ctx* ctx, int world, int rank, int lanes, uint8_t* handle_out) {
    if (!ctx || !handle_out || world < 1 || world > 8 || rank < 0 || rank >= world || lanes < 1 || lanes > 64) return fail(ctx, BGLS_ERR_ARG, "bad argument");
    Exchange& x = ctx->xch;
    if (x.local) return fail(ctx, BGLS_ERR_ARG, "exchange already created");
    CU(cudaSetDevice(ctx->device));
    x.world = world; x.rank = rank; x.lanes = lanes;
    const size_t bytes = xch_flag_base(x) + (size_t)lanes * 2 * world * sizeof(unsigned long long);
    CU(cudaMalloc((void**)&x.local, bytes));
    CU(cudaMemset(x.local, 0, bytes));
    CU(cudaMalloc((void**)&x.d_err, sizeof(int)));
    CU(cudaMemset(x.d_err, 0, sizeof(int)));
    CU(cudaDeviceSynchronize());
    cudaIpcMemHandle_t h;
    CU(cudaIpcGetMemHandle(&h, x.local));
    static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
    memcpy(handle_out, &h, sizeof(h));
    x.peer.assign(world, nullptr);
    x.peer[rank] = x.local;
    x.ready = world == 1;
    return BGLS_OK;
}
int bgls_exchange_connect(bgls_ctx* ctx, int peer_rank, const uint8_t* handle) {
    if (!ctx || !handle || !ctx->xch.local || peer_rank < 0 || peer_rank >= ctx->xch.world) return fail(ctx, BGLS_ERR_ARG, "bad argument");
    Exchange& x = ctx->xch;
    CU(cudaSetDevice(ctx->device));
    if (peer_rank != x.rank && !x.peer[peer_rank]) {
        cudaIpcMemHandle_t h;
        memcpy(&h, handle, sizeof(h));
        void* p = nullptr;
        CU(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        x.peer[peer_rank] = (uint8_t*)p;
    }
    bool all = true;
    for (uint8_t* q : x.peer) all = all && q != nullptr;
    x.ready = all;
    return BGLS_OK;
}
int bgls_exchange_error(bgls_ctx* ctx, int* err) {
    if (!ctx || !err || !ctx->xch.d_err) return fail(ctx, BGLS_ERR_ARG, "bad argument");
    CU(cudaSetDevice(ctx->device));
    CU(cudaMemcpy(err, ctx->xch.d_err, sizeof(int), cudaMemcpyDeviceToHost));
    if (*err) CU(cudaMemset(ctx->xch.d_err, 0, sizeof(int)));   // reported once: a timeout does not poison later epochs
    return BGLS_OK;
}
int bgls_miller_product_exchange_dev(bgls_ctx* ctx, int curve, const void* d_g1, const void* d_g2, size_t n, int lane,
                                     uint64_t epoch, void* stream) {
    if (!ctx || !curve_ok(curve) || (n && (!d_g1 || !d_g2)) || !ctx->xch.ready || lane < 0 || lane >= ctx->xch.lanes || epoch == 0)
        return fail(ctx, BGLS_ERR_ARG, "bad argument / exchange not connected");
    Exchange& x = ctx->xch;
    SlotLock sl(ctx, (cudaStream_t)stream);
    CU(cudaSetDevice(ctx->device));
    const size_t F = fp_bytes(curve), rec = 12 * F, bw = pairing_work_bytes(ctx, curve, n);
    void* work;
    int rc = dev_work(ctx, sl.s, bw + align_up(rec + 16), &work);
    if (rc) return rc;
    uint8_t* partial = (uint8_t*)work + bw;
    PeerPtrs pp{};
    SvPeers sp{};
    const int parity = (int)(epoch & 1);
    for (int r = 0; r < x.world; r++) {
        sp.slot[r] = pp.slot[r] = x.peer[r] + xch_rec_off(x, lane, parity, x.rank, rec);
        sp.flag[r] = pp.flag[r] = (unsigned long long*)(x.peer[r] + xch_flag_off(x, lane, parity, x.rank));
    }
    sp.world = x.world;
    sp.nbytes = (int)rec;
    sp.epoch = (unsigned long long)epoch;
    bool fused = false;
    rc = pairing_dev(ctx, curve, d_g1, d_g2, n, 0, partial, nullptr, work, (cudaStream_t)stream, &sp, &fused);
    if (rc) return rc;
    if (!fused) {
        // broadcast over peer memory: one tiny kernel, payload then flag (the slot engine does it inside its own launch)
        k_exchange_send_t<PeerPtrs><<<1, 256, 0, (cudaStream_t)stream>>>(partial, (int)rec, pp, x.world, (unsigned long long)epoch);
        ctx->launches++;
    }
    CU(cudaGetLastError());
    return BGLS_OK;
}
int bgls_final_exp_exchanged_dev(bgls_ctx* ctx, int curve, int lane, uint64_t epoch, void* d_out_gt, void* d_is_identity, void* stream) {
    if (!ctx || !curve_ok(curve) || !d_out_gt || !ctx->xch.ready || lane < 0 || lane >= ctx->xch.lanes || epoch == 0)
        return fail(ctx, BGLS_ERR_ARG, "bad argument / exchange not connected");
    Exchange& x = ctx->xch;
    SlotLock sl(ctx, (cudaStream_t)stream);
    CU(cudaSetDevice(ctx->device));
    const size_t F = fp_bytes(curve), rec = 12 * F;
    const int parity = (int)(epoch & 1);
    void* work;
    int rc = dev_work(ctx, sl.s, mach_work_for(curve, x.world), &work);
    if (rc) return rc;
    const unsigned long long* flags = (const unsigned long long*)(x.local + xch_flag_off(x, lane, parity, 0));
    const XchWait xw{flags, (unsigned long long)epoch, x.d_err};
    // throughput regime: wait, product and final exponentiation are ONE launch (k_slot_finish_bytes); otherwise wait kernel,
    // the machine's import / tree / finish, guard kernel
    const bool slot_path = d_is_identity && !ctx->thread_engine && (size_t)x.world <= SLOT_FANIN &&
                           (ctx->engine == ENGINE_SLOT || (ctx->engine == ENGINE_AUTO && busy_estimate(ctx) >= SLOT_MIN_BUSY));
    if (!slot_path) {
        k_exchange_wait<<<1, 32, 0, (cudaStream_t)stream>>>(flags, x.world, (unsigned long long)epoch, x.d_err);
        ctx->launches++;
    }
    bool fused = false;
    rc = finish_bytes_dev(ctx, curve, x.local + xch_rec_off(x, lane, parity, 0, rec), x.world, 1, d_out_gt, d_is_identity, work, (cudaStream_t)stream,
                          slot_path ? &xw : nullptr, &fused);
    if (rc) return rc;
    if (slot_path && !fused) return fail(ctx, BGLS_ERR_ARG, "exchange: engine choice changed between the wait and the finish");
    if (!fused) {
        k_exchange_guard<<<1, 1, 0, (cudaStream_t)stream>>>(x.d_err, (int*)d_is_identity);   // timed-out wait -> verdict false
        ctx->launches++;
    }
    CU(cudaGetLastError());
    return BGLS_OK;
}
int bgls_final_exp_product_dev(bgls_ctx* ctx, int curve, const void* d_partials, size_t k, void* d_out_gt,
                               void* d_is_identity, void* stream) {
    if (!ctx || !curve_ok(curve) || !d_out_gt || (k && !d_partials)) return fail(ctx, BGLS_ERR_ARG, "bad argument");
    SlotLock sl(ctx, (cudaStream_t)stream);
    CU(cudaSetDevice(ctx->device));
    void* work;
    int rc = dev_work(ctx, sl.s, mach_work_for(curve, k), &work);
    if (rc) return rc;
    return finish_bytes_dev(ctx, curve, d_partials, k, 1, d_out_gt, d_is_identity, work, (cudaStream_t)stream);
}
int bgls_aggregate_points_dev(bgls_ctx* ctx, int curve, int group, const void* d_pts, size_t n, void* d_out, void* stream) {
    if (!ctx || !curve_ok(curve) || (group != 1 && group != 2) || !d_pts || !d_out || n == 0) return fail(ctx, BGLS_ERR_ARG, "bad argument");
    SlotLock sl(ctx, (cudaStream_t)stream);
    CU(cudaSetDevice(ctx->device));
    void* work;
    int rc = dev_work(ctx, sl.s, agg_work_bytes(curve, group, n), &work);
    if (rc) return rc;
    return aggregate_dev(ctx, curve, group, d_pts, n, d_out, work, (cudaStream_t)stream);
}
int bgls_scale_points_dev(bgls_ctx* ctx, int curve, int group, const void* d_pts, const void* d_scalars, size_t n,
                          void* d_out, void* stream) {
    if (!ctx || !curve_ok(curve) || (group != 1 && group != 2) || (n && (!d_pts || !d_scalars || !d_out))) return fail(ctx, BGLS_ERR_ARG, "bad argument");
    CU(cudaSetDevice(ctx->device));
    return scale_dev(ctx, curve, group, d_pts, d_scalars, n, d_out, (cudaStream_t)stream);
}
int bgls_pairing_check_batch_dev(bgls_ctx* ctx, int curve, const void* d_g1, const void* d_g2, const void* d_offsets,
                                 size_t nbatch, size_t total_pairs, void* d_out_ok, void* stream) {
    if (!ctx || !curve_ok(curve) || !d_offsets || (nbatch && !d_out_ok) || (total_pairs && (!d_g1 || !d_g2))) return fail(ctx, BGLS_ERR_ARG, "bad argument");
    SlotLock sl(ctx, (cudaStream_t)stream);
    CU(cudaSetDevice(ctx->device));
    void* work;
    int rc = dev_work(ctx, sl.s, std::max({align_up(total_pairs * fp12_dev_bytes(curve)), mach_work_for(curve, total_pairs + nbatch), slot_batch_work_bytes(curve, nbatch, total_pairs)}), &work);
    if (rc) return rc;
    return batch_dev(ctx, curve, d_g1, d_g2, d_offsets, nbatch, total_pairs, d_out_ok, work, (cudaStream_t)stream);
}

}  // extern "C"
