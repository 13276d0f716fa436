// Device kernels of the dot-product machine (see machine.cuh / tools/gen_machine.py).
//   k_mach_miller    K1: one Miller loop per 16-lane group, two groups per warp
//   k_mach_reduce    K2: Fp12 product tree (chunked or segmented), one group per output
//   k_mach_finish    K3: final exponentiation (or plain export) + wire bytes + identity flag
//   k_mach_import    wire bytes -> machine form (for bgls_final_exp_product)
// Internal Fp12 values travel between kernels as [12][L] uint32 limbs (Montgomery, unsaturated).
#pragma once
#include <cuda_runtime.h>

#include "machine.cuh"
#include "machine_tables.cuh"

namespace bgls {

constexpr int MWPB = 4;                 // warps per block
constexpr int MP_MAXW = 12;             // k_mach_miller32: pairings (warps) per block, at most
constexpr int MT_MAXW = 16;             // k_mach_tree32: values (warps) per block, at most
constexpr int MFIN_THREADS = 256;       // k_mach_finish: one working warp, the others only stage the tables
constexpr int MGPB = MWPB * 2;          // groups per block

// shared memory layout of every machine kernel (32-bit words):
//   [ phase headers | phase records (u16 x REC x LANES per phase) | group slot files (NS slots each) ]
// Keeping the schedule tables on chip removes every global load from the phase loop; every group file
// ends with its own copy of the constants so that all operands share one addressing form.
template <class M> __host__ __device__ constexpr size_t mach_tab_words() {
    return (size_t)M::NPHASE + ((size_t)M::NPHASE * M::LANES * M::REC + 1) / 2;
}
template <class M> constexpr size_t mach_smem_bytes(int groups) {
    return (mach_tab_words<M>() + (size_t)groups * M::NS * M::L) * sizeof(uint32_t);
}

// copies the tables into shared memory and the constants into the `groups` group files; returns the group
// area and rewrites `tb` to the on-chip copies
template <class M> __device__ __forceinline__ uint32_t* mach_stage_tables(uint32_t* sm, MachTables& tb, int groups) {
    uint32_t* hdr = sm;
    uint32_t* rec = hdr + M::NPHASE;
    uint32_t* gbase = sm + mach_tab_words<M>();
    for (int idx = threadIdx.x; idx < M::NPHASE; idx += blockDim.x) hdr[idx] = tb.hdr[idx];
    const uint32_t* rsrc = (const uint32_t*)tb.rec;   // NPHASE*LANES*REC u16 = even number of u16
    constexpr int RW = M::NPHASE * M::LANES * M::REC / 2;
    for (int idx = threadIdx.x; idx < RW; idx += blockDim.x) rec[idx] = rsrc[idx];
    constexpr int CW = M::NCONST * M::L;
    for (int idx = threadIdx.x; idx < groups * CW; idx += blockDim.x) {
        const int g = idx / CW, r = idx % CW, c = r / M::L, i = r % M::L;
        gbase[(size_t)g * M::NS * M::L + i * M::NS + M::NSG + c] = tb.consts[r];
    }
    __syncthreads();
    tb.hdr = hdr;
    tb.rec = (const uint16_t*)rec;
    return gbase;
}

template <class M> __device__ __forceinline__ void mach_run(const MachView<M>& mv, const MachTables& tb,
                                                            const uint16_t* __restrict__ prog, int len, int gl,
                                                            unsigned wmask = 0xFFFFFFFFu) {
    uint32_t ph = __ldg(prog);
    for (int pc = 0; pc < len; pc++) {
        const uint32_t nxt = __ldg(prog + (pc + 1 < len ? pc + 1 : pc));  // prefetch: hides the global latency
        mach_phase_lane<M>(mv, tb, ph, gl);
        __syncwarp(wmask);
        ph = nxt;
    }
}

// ---------------------------------------------------------------- K1
// GPW = groups per warp.  GPW = 2 packs two pairings into each warp (throughput mode).  GPW = 1 leaves
// lanes 16..31 idle but doubles the number of resident warps: one warp per SM sub-partition only
// reaches about half of the IMAD.WIDE issue rate (ncu: issue active 27 %, stall_wait 2.2), so small
// products (the 1025-pair aggregate verify) run faster with more, half-empty warps.
template <class M, int GPW>
__global__ void __launch_bounds__(MWPB * 32) k_mach_miller(MachTables tb, const uint16_t* __restrict__ prog, int plen,
                                                          const uint8_t* __restrict__ g1, const uint8_t* __restrict__ g2,
                                                          size_t n, uint32_t* __restrict__ vals) {
    extern __shared__ uint32_t sm[];
    constexpr int L = M::L, FB = M::FP_BYTES;
    constexpr int GPB = MWPB * GPW;  // groups per block
    uint32_t* gbase = mach_stage_tables<M>(sm, tb, GPB);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (GPW == 1 && lane >= MG) return;
    const int g = GPW == 2 ? (threadIdx.x >> 4) : warp, gl = lane & 15;
    const size_t pair = (size_t)blockIdx.x * GPB + g;
    const size_t warp_first = (size_t)blockIdx.x * GPB + (GPW == 2 ? (g & ~1) : g);
    if (warp_first >= n) return;  // whole warp idle
    MachView<M> mv;
    mv.gs = gbase + (size_t)g * M::NS * L;
    const bool active = pair < n;
    // inputs: lanes 0..5 of the group convert one coordinate each
    bool zero = true, flag = false;
    if (active && gl < 6) {
        const uint8_t* src = gl < 2 ? g1 + pair * 2 * FB + gl * FB : g2 + pair * 4 * FB + (gl - 2) * FB;
        const int slot = gl == 0 ? M::IN_XP : gl == 1 ? M::IN_YP : gl == 2 ? M::IN_XQY : gl == 3 ? M::IN_XQX : gl == 4 ? M::IN_YQY : M::IN_YQX;
        uint32_t v[L];
        mach_limbs_from_be<M>(v, src);
        uint32_t any = 0;
#pragma unroll
        for (int i = 0; i < L; i++) any |= v[i];
        zero = any == 0;
        flag = FB == 48 && (gl == 0 || gl == 2) && (src[0] & 0x40);
        mach_store<M>(mv, slot, v);
    }
    const unsigned wmask = GPW == 2 ? 0xFFFFFFFFu : 0xFFFFu;
    const unsigned zb = (__ballot_sync(wmask, zero) >> (lane & 16)) & 0xFFFFu;
    const unsigned fb = (__ballot_sync(wmask, flag) >> (lane & 16)) & 0xFFFFu;
    const bool inf = ((zb & 0x3u) == 0x3u) || ((zb & 0x3Cu) == 0x3Cu) || (fb & 0x5u);
    __syncwarp(wmask);
    mach_run<M>(mv, tb, prog, plen, gl, wmask);
    if (active && gl < 12) {
        uint32_t v[L];
        mach_load<M>(v, mv, inf ? (gl == 0 ? M::ONE : M::ZERO) : M::FA0 + gl);
        uint32_t* o = vals + (pair * 12 + gl) * L;
#pragma unroll
        for (int i = 0; i < L; i++) o[i] = v[i];
    }
}

// ---------------------------------------------------------------- K1p
// Latency path: one pairing per warp running the pipelined program of the signed 32-lane slot file P
// (tools/gen_machine.py: build_miller_p).  blockDim.x / 32 pairings per block, chosen by the host so that the
// grid covers all SMs with as many resident warps per SM sub-partition as the product offers.
// In-block binary product tree over the FA registers of the first `wact` group files of the block (one warp each,
// P::MULACC multiplies in place): log2 levels, each one Fp12 product deep.  Every warp of the block must call it.
// Afterwards FA of group 0 holds the product.
template <class P>
__device__ __forceinline__ void mach_block_tree(uint32_t* gbase, const MachTables& tb, const uint16_t* __restrict__ prog_mul,
                                                int wact, int warp, int lane) {
    constexpr int L = P::L;
    MachView<P> mv;
    mv.gs = gbase + (size_t)warp * P::NS * L;
    for (int stride = 1; stride < wact; stride <<= 1) {
        __syncthreads();   // the partner's product of the previous level is complete
        if ((warp & (2 * stride - 1)) == 0 && warp + stride < wact) {
            const uint32_t* pg = gbase + (size_t)(warp + stride) * P::NS * L;
            for (int idx = lane; idx < 12 * L; idx += 32) {
                const int k = idx / L, i = idx % L;
                mv.gs[i * P::NS + P::GB0 + k] = pg[i * P::NS + P::FA0 + k];
            }
            __syncwarp();
            mach_run<P>(mv, tb, prog_mul, P::MULACC_LEN, lane);
        }
    }
}

// launch bounds: for the 10-limb curve 12 warps per block and two blocks per SM keep the kernel at 80 registers (81
// without them: the allocation granule then makes a 12-warp block take more than half of the register file); the
// 14-limb curve needs 98 registers and is left alone (forcing 80 spills: measured 16 % slower)
template <class P>
__global__ void __launch_bounds__(MP_MAXW * 32, P::L <= 10 ? 2 : 1) k_mach_miller32(MachTables tb, const uint16_t* __restrict__ prog, int plen,
                                                      const uint16_t* __restrict__ prog_mul, int fuse,
                                                      const uint8_t* __restrict__ g1, const uint8_t* __restrict__ g2,
                                                      size_t n, uint32_t* __restrict__ vals) {
    // fuse != 0: vals[block] = product of the block's Miller values; fuse == 0: vals[pair] = Miller value of the pair
    extern __shared__ uint32_t sm[];
    constexpr int L = P::L, FB = P::FP_BYTES;
    static_assert(P::LANES == 32, "one warp per pairing");
    const int wpb = blockDim.x >> 5;
    uint32_t* gbase = mach_stage_tables<P>(sm, tb, wpb);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t first = (size_t)blockIdx.x * wpb;
    const size_t pair = first + warp;
    const int wact = (int)(n - first < (size_t)wpb ? n - first : (size_t)wpb);   // pairings of this block
    MachView<P> mv;
    mv.gs = gbase + (size_t)warp * P::NS * L;
    if (pair < n) {
        bool zero = true, flag = false;
        if (lane < 6) {   // lanes 0..5 convert one coordinate each
            const uint8_t* src = lane < 2 ? g1 + pair * 2 * FB + lane * FB : g2 + pair * 4 * FB + (lane - 2) * FB;
            const int slot = lane == 0 ? P::IN_XP : lane == 1 ? P::IN_YP : lane == 2 ? P::IN_XQY : lane == 3 ? P::IN_XQX : lane == 4 ? P::IN_YQY : P::IN_YQX;
            uint32_t v[L];
            mach_limbs_from_be<P>(v, src);
            uint32_t any = 0;
#pragma unroll
            for (int i = 0; i < L; i++) any |= v[i];
            zero = any == 0;
            flag = FB == 48 && (lane == 0 || lane == 2) && (src[0] & 0x40);
            mach_store<P>(mv, slot, v);
        }
        const unsigned zb = __ballot_sync(0xFFFFFFFFu, zero) & 0x3Fu;
        const unsigned fb = __ballot_sync(0xFFFFFFFFu, flag) & 0x3Fu;
        const bool inf = ((zb & 0x3u) == 0x3u) || ((zb & 0x3Cu) == 0x3Cu) || (fb & 0x5u);
        __syncwarp();
        mach_run<P>(mv, tb, prog, plen, lane);
        if (inf && lane < 12) {   // e(P, Q) = 1 when either point is the point at infinity
            uint32_t v[L];
            mach_load<P>(v, mv, lane == 0 ? P::ONE : P::ZERO);
            mach_store<P>(mv, P::FA0 + lane, v);
        }
        __syncwarp();
    }
    if (fuse) mach_block_tree<P>(gbase, tb, prog_mul, wact, warp, lane);   // one output per block
    if ((fuse ? warp == 0 : pair < n) && lane < 12) {
        uint32_t v[L];
        mach_load<P>(v, mv, P::FA0 + lane);
        uint32_t* o = vals + ((fuse ? (size_t)blockIdx.x : pair) * 12 + lane) * L;
#pragma unroll
        for (int i = 0; i < L; i++) o[i] = v[i];
    }
}

// ---------------------------------------------------------------- K2p
// One level of the product tree: every block multiplies up to blockDim.x / 32 values (machine form) into one.
template <class P>
__global__ void __launch_bounds__(MT_MAXW * 32) k_mach_tree32(MachTables tb, const uint16_t* __restrict__ prog_mul,
                                                    const uint32_t* __restrict__ in, size_t n, uint32_t* __restrict__ out) {
    extern __shared__ uint32_t sm[];
    constexpr int L = P::L;
    const int wpb = blockDim.x >> 5;
    uint32_t* gbase = mach_stage_tables<P>(sm, tb, wpb);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t first = (size_t)blockIdx.x * wpb;
    const int wact = (int)(n - first < (size_t)wpb ? n - first : (size_t)wpb);
    MachView<P> mv;
    mv.gs = gbase + (size_t)warp * P::NS * L;
    if (warp < wact) {
        const uint32_t* src = in + (first + warp) * 12 * L;
        for (int idx = lane; idx < 12 * L; idx += 32) mv.gs[(idx % L) * P::NS + P::FA0 + idx / L] = src[idx];
        __syncwarp();
    }
    mach_block_tree<P>(gbase, tb, prog_mul, wact, warp, lane);
    if (warp == 0 && lane < 12) {
        uint32_t v[L];
        mach_load<P>(v, mv, P::FA0 + lane);
        uint32_t* o = out + ((size_t)blockIdx.x * 12 + lane) * L;
#pragma unroll
        for (int i = 0; i < L; i++) o[i] = v[i];
    }
}

// ---------------------------------------------------------------- K2
// output g = product of inputs [lo, hi): lo/hi from `seg` (nout+1 offsets) when given, else chunks of `chunk`
template <class M>
__global__ void __launch_bounds__(MWPB * 32) k_mach_reduce(MachTables tb, const uint16_t* __restrict__ prog_ab,
                                                          const uint16_t* __restrict__ prog_ba, const uint32_t* __restrict__ in,
                                                          size_t n_in, size_t chunk, const unsigned long long* __restrict__ seg,
                                                          size_t n_out, uint32_t* __restrict__ out) {
    extern __shared__ uint32_t sm[];
    constexpr int L = M::L;
    uint32_t* gbase = mach_stage_tables<M>(sm, tb, MGPB);
    const int g = threadIdx.x >> 4, gl = threadIdx.x & 15;
    const size_t o = (size_t)blockIdx.x * MGPB + g;
    const size_t warp_first = (size_t)blockIdx.x * MGPB + (g & ~1);
    if (warp_first >= n_out) return;
    MachView<M> mv;
    mv.gs = gbase + (size_t)g * M::NS * L;
    const bool active = o < n_out;
    size_t lo = 0, hi = 0;
    if (active) {
        if (seg) { lo = seg[o]; hi = seg[o + 1]; }
        else { lo = o * chunk; hi = lo + chunk < n_in ? lo + chunk : n_in; }
    }
    // the two groups of a warp walk in lock step: iterate to the longer of the two segments
    const size_t len = hi - lo;
    const size_t other = __shfl_xor_sync(0xFFFFFFFFu, (unsigned long long)len, 16);
    const size_t steps = len > other ? len : other;
    bool in_a = true;
    if (gl < 12) {
        uint32_t v[L];
        if (len > 0) {
            const uint32_t* src = in + (lo * 12 + gl) * L;
#pragma unroll
            for (int i = 0; i < L; i++) v[i] = src[i];
        } else {
            mach_load<M>(v, mv, gl == 0 ? M::ONE : M::ZERO);
        }
        mach_store<M>(mv, M::FA0 + gl, v);
    }
    __syncwarp();
    for (size_t s = 1; s < steps; s++) {
        if (gl < 12) {
            uint32_t v[L];
            if (s < len) {
                const uint32_t* src = in + ((lo + s) * 12 + gl) * L;
#pragma unroll
                for (int i = 0; i < L; i++) v[i] = src[i];
            } else {
                mach_load<M>(v, mv, gl == 0 ? M::ONE : M::ZERO);  // multiply by one: keeps the warp uniform
            }
            mach_store<M>(mv, M::GB0 + gl, v);
        }
        __syncwarp();
        mach_run<M>(mv, tb, in_a ? prog_ab : prog_ba, M::MUL_AB_LEN, gl);
        in_a = !in_a;
    }
    if (active && gl < 12) {
        uint32_t v[L];
        mach_load<M>(v, mv, (in_a ? M::FA0 : M::FB0) + gl);
        uint32_t* dst = out + (o * 12 + gl) * L;
#pragma unroll
        for (int i = 0; i < L; i++) dst[i] = v[i];
    }
}

// ---------------------------------------------------------------- K3
// one 32-lane group (one warp, one block) per value: FINALEXP (or EXPORT) then canonical wire bytes;
// flags[o] = 1 iff the result is 1.  The F slot file splits every long dot over two lanes.
template <class F, class MIN>
__global__ void __launch_bounds__(MFIN_THREADS) k_mach_finish(MachTables tb, const uint16_t* __restrict__ prog, int plen,
                                                   const uint32_t* __restrict__ in, size_t n, uint8_t* __restrict__ out_gt,
                                                   int* __restrict__ flags32, uint8_t* __restrict__ flags8) {
    extern __shared__ uint32_t sm[];
    constexpr int L = F::L, FB = F::FP_BYTES;
    static_assert(F::L == MIN::L, "limb layout mismatch between slot files");
    static_assert(F::LANES == 32, "the final-exponentiation slot file is scheduled for 32-lane groups");
    uint32_t* gbase = mach_stage_tables<F>(sm, tb, 1);
    if (threadIdx.x >= 32) return;   // the extra warps only help staging the tables
    const int gl = threadIdx.x;
    const size_t o = blockIdx.x;
    MachView<F> mv;
    mv.gs = gbase;
    if (gl < 12) {
        uint32_t v[L];
        const uint32_t* src = in + (o * 12 + gl) * L;
#pragma unroll
        for (int i = 0; i < L; i++) v[i] = src[i];
        mach_store<F>(mv, F::FA0 + gl, v);
    }
    __syncwarp();
    mach_run<F>(mv, tb, prog, plen, gl);
    bool ok = true;
    if (gl < 12) {
        uint32_t v[L];
        mach_load<F>(v, mv, F::OUT0 + gl);
        const int k = gl >> 1, part = gl & 1;                       // coefficient of w^k, 0 = re, 1 = im
        const int pos = k == 5 ? 0 : k == 3 ? 1 : k == 1 ? 2 : k == 4 ? 3 : k == 2 ? 4 : 5;  // GT order 5,3,1,4,2,0
        bool z, one;
        uint8_t bytes[FB];
        mach_canon_be<F>(bytes, v, &z, &one);
        if (out_gt) {
            uint8_t* dst = out_gt + o * 12 * FB + (size_t)(2 * pos + (part ? 0 : 1)) * FB;
#pragma unroll
            for (int i = 0; i < FB; i++) dst[i] = bytes[i];
        }
        ok = (gl == 0) ? one : z;
    }
    const unsigned okb = __ballot_sync(0xFFFFFFFFu, ok);
    if (gl == 0) {
        const int is_one = (okb & 0xFFFu) == 0xFFFu;
        if (flags32) flags32[o] = is_one;
        if (flags8) flags8[o] = (uint8_t)is_one;
    }
}

// Low-footprint final exponentiation for the throughput regime: the schedule tables stay in global memory (L2 / L1 hits)
// and only the group file lives in shared memory (19 / 26 KB instead of 148 / 190 KB).  It runs as the tail of the slot
// engine's Miller kernel (one launch per product: a dependent launch would stall the hardware queue its stream shares with
// other streams, and a block that needs most of an SM's shared memory is starved while the GPU is full of small Miller
// blocks) or as a one-warp kernel of its own.  About 0.1 ms slower than the staged kernel when it runs alone.
struct MachFinishArgs {
    MachTables tb;
    const uint16_t* prog;
    int plen;
    uint8_t* out_gt;      // 12 F bytes (may be null)
    int* flag32;          // is-identity flag (may be null)
    uint8_t* flag8;       // is-identity flag, byte form (batches; may be null)
};
// one warp; `sm`: at least F::NS * F::L words of shared memory; `in`: the value in machine form ([12][L] limbs)
template <class F>
__device__ __forceinline__ void mach_finish_warp(uint32_t* sm, const MachFinishArgs& a, const uint32_t* in) {
    constexpr int L = F::L, FB = F::FP_BYTES;
    static_assert(F::LANES == 32, "the final-exponentiation slot file is scheduled for 32-lane groups");
    uint32_t* gbase = sm;
    constexpr int CW = F::NCONST * L;
    const int gl = threadIdx.x & 31;
    for (int idx = gl; idx < CW; idx += 32) gbase[(idx % L) * F::NS + F::NSG + idx / L] = a.tb.consts[idx];
    MachView<F> mv;
    mv.gs = gbase;
    if (gl < 12) {
        uint32_t v[L];
        const uint32_t* src = in + gl * L;
#pragma unroll
        for (int i = 0; i < L; i++) v[i] = __ldcg(src + i);
        mach_store<F>(mv, F::FA0 + gl, v);
    }
    __syncwarp();
    mach_run<F>(mv, a.tb, a.prog, a.plen, gl);
    bool ok = true;
    if (gl < 12) {
        uint32_t v[L];
        mach_load<F>(v, mv, F::OUT0 + gl);
        const int k = gl >> 1, part = gl & 1;                       // coefficient of w^k, 0 = re, 1 = im
        const int pos = k == 5 ? 0 : k == 3 ? 1 : k == 1 ? 2 : k == 4 ? 3 : k == 2 ? 4 : 5;  // GT order 5,3,1,4,2,0
        bool z, one;
        uint8_t bytes[FB];
        mach_canon_be<F>(bytes, v, &z, &one);
        if (a.out_gt) {
            uint8_t* dst = a.out_gt + (size_t)(2 * pos + (part ? 0 : 1)) * FB;
#pragma unroll
            for (int i = 0; i < FB; i++) dst[i] = bytes[i];
        }
        ok = (gl == 0) ? one : z;
    }
    const unsigned okb = __ballot_sync(0xFFFFFFFFu, ok);
    if (gl == 0 && a.flag32) *a.flag32 = (okb & 0xFFFu) == 0xFFFu;
    if (gl == 0 && a.flag8) *a.flag8 = (uint8_t)((okb & 0xFFFu) == 0xFFFu);
}
template <class F> struct MachFinisher {   // tail of k_slot_miller (slotvm.cuh)
    using Args = MachFinishArgs;
    static constexpr size_t SMEM_BYTES = (size_t)F::NS * F::L * sizeof(uint32_t);
    static constexpr bool HAS_TAIL = true;
    static constexpr int MACH_L = F::L;
    __device__ __forceinline__ static void run(uint32_t* sm, const Args& a, const uint32_t* in) { mach_finish_warp<F>(sm, a, in); }
    __device__ __forceinline__ static const uint8_t* wire_bytes(const Args& a) { return a.out_gt; }
    __device__ __forceinline__ static void force_false(const Args& a) {
        if (a.flag32) *a.flag32 = 0;
        if (a.flag8) *a.flag8 = 0;
    }
    __device__ __forceinline__ static Args for_product(Args a, uint8_t* flags8, size_t c) {   // product c of a batch
        a.out_gt = nullptr;
        a.flag32 = nullptr;
        a.flag8 = flags8 + c;
        return a;
    }
};
template <class F>
__global__ void __launch_bounds__(32) k_mach_finish_lean(MachFinishArgs a, const uint32_t* __restrict__ in) {
    extern __shared__ uint32_t sm[];
    mach_finish_warp<F>(sm, a, in);
}

// ---------------------------------------------------------------- import: wire GT bytes -> machine form
template <class M>
__global__ void __launch_bounds__(MWPB * 32) k_mach_import(MachTables tb, const uint16_t* __restrict__ prog, int plen,
                                                          const uint8_t* __restrict__ in, size_t n, uint32_t* __restrict__ vals) {
    extern __shared__ uint32_t sm[];
    constexpr int L = M::L, FB = M::FP_BYTES;
    uint32_t* gbase = mach_stage_tables<M>(sm, tb, MGPB);
    const int g = threadIdx.x >> 4, gl = threadIdx.x & 15;
    const size_t o = (size_t)blockIdx.x * MGPB + g;
    const size_t warp_first = (size_t)blockIdx.x * MGPB + (g & ~1);
    if (warp_first >= n) return;
    MachView<M> mv;
    mv.gs = gbase + (size_t)g * M::NS * L;
    const bool active = o < n;
    if (gl < 12) {
        uint32_t v[L];
        if (active) {
            const int k = gl >> 1, part = gl & 1;
            const int pos = k == 5 ? 0 : k == 3 ? 1 : k == 1 ? 2 : k == 4 ? 3 : k == 2 ? 4 : 5;
            mach_limbs_from_be<M>(v, in + o * 12 * FB + (size_t)(2 * pos + (part ? 0 : 1)) * FB);
        } else {
#pragma unroll
            for (int i = 0; i < L; i++) v[i] = 0;
        }
        mach_store<M>(mv, M::RAWF0 + gl, v);
    }
    __syncwarp();
    mach_run<M>(mv, tb, prog, plen, gl);
    if (active && gl < 12) {
        uint32_t v[L];
        mach_load<M>(v, mv, M::FA0 + gl);
        uint32_t* dst = vals + (o * 12 + gl) * L;
#pragma unroll
        for (int i = 0; i < L; i++) dst[i] = v[i];
    }
}

}  // namespace bgls
