// Slot engine: Miller loops of the pairing product as straight-line programs of Fp2 operations (tools/gen_slotvm.py,
// slotvm_tables.cuh) over a per-pair file of Fp2 slots in shared memory, on saturated 32-bit limbs with lazy
// reduction (sat.cuh).
//
// Replaces the Miller half of every `Pair` of concurrentPairingProduct (/root/reference/curves/curve.go:125-170,
// 217-223; altbn128.go:130-141; bls12_381.go:228-236) and the first levels of its GT product tree
// (curve.go:141-169).  A pair is owned by G lanes (G = 1, 2 or 4): every round of a program holds one operation per
// lane; the lanes of a pair exchange values only through the slot file and synchronise with __syncwarp().
//
// Slot file layout (uint4 units): slot s, limb group w (4 limbs), pair q  ->  (s * W4 + w) * NPB + q, i.e. the lanes of a
// warp read consecutive 16-byte words (conflict free, LDS.128 / STS.128).  An Fp2 slot is c0 (N limbs) then c1.
// Slots >= 256 are the block-shared constants (canonical Montgomery form), laid out contiguously.
//
// Host build (tests/host_emul): the same interpreter runs sequentially over the lanes.
#pragma once
#include <cstdint>

#include "sat.cuh"
#include "slotvm_tables.cuh"

namespace bgls {

enum : uint32_t { SV_NOP = 0, SV_MUL, SV_SQR, SV_ADD, SV_SUB, SV_XI, SV_HALF, SV_CONJ, SV_NEG, SV_COPY, SV_SEL0, SV_SEL1 };
constexpr uint32_t SV_CONST0 = 256;

struct alignas(16) SvU4 { uint32_t x, y, z, w; };   // 16-byte unit of the slot file (uint4 on the device)

template <class C, int NPB> struct SlotFile {
    static constexpr int N = C::N, W4 = 2 * N / 4;
    SvU4* slots;          // group slots
    const SvU4* consts;   // constants, [c][W4]
    int q;                // group index inside the block
    uint32_t flags;       // bit j: pair j of the group has a point at infinity (SEL0 / SEL1)

    HD F2<C> load(uint32_t s) const {
        const SvU4* p;
        int stride;
        if (s >= SV_CONST0) { p = consts + (s - SV_CONST0) * W4; stride = 1; }
        else { p = slots + (size_t)s * W4 * NPB + q; stride = NPB; }
        F2<C> r;
#pragma unroll
        for (int w = 0; w < W4; w++) {
            const SvU4 t = p[w * stride];
            uint32_t* v = w < N / 4 ? r.c0.v + 4 * w : r.c1.v + 4 * (w - N / 4);
            v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
        }
        return r;
    }
    HD void store(uint32_t s, const F2<C>& a) const {
        SvU4* p = slots + (size_t)s * W4 * NPB + q;
#pragma unroll
        for (int w = 0; w < W4; w++) {
            const uint32_t* v = w < N / 4 ? a.c0.v + 4 * w : a.c1.v + 4 * (w - N / 4);
            SvU4 t;
            t.x = v[0]; t.y = v[1]; t.z = v[2]; t.w = v[3];
            p[w * NPB] = t;
        }
    }
    // one Fp component (half = 0: c0, 1: c1) of a pair slot
    HD void store_fp(uint32_t s, int half, const LN<N>& a) const {
        SvU4* p = slots + ((size_t)s * W4 + half * (N / 4)) * NPB + q;
#pragma unroll
        for (int w = 0; w < N / 4; w++) {
            SvU4 t;
            t.x = a.v[4 * w]; t.y = a.v[4 * w + 1]; t.z = a.v[4 * w + 2]; t.w = a.v[4 * w + 3];
            p[w * NPB] = t;
        }
    }
    HD LN<N> load_fp(uint32_t s, int half) const {
        const SvU4* p = slots + ((size_t)s * W4 + half * (N / 4)) * NPB + q;
        LN<N> r;
#pragma unroll
        for (int w = 0; w < N / 4; w++) {
            const SvU4 t = p[w * NPB];
            r.v[4 * w] = t.x; r.v[4 * w + 1] = t.y; r.v[4 * w + 2] = t.z; r.v[4 * w + 3] = t.w;
        }
        return r;
    }
};

// one operation of one lane
template <class C, int NPB> HD void sv_exec(const SlotFile<C, NPB>& sf, uint32_t op) {
    const uint32_t kind = op & 31u, d = (op >> 5) & 511u, sa = (op >> 14) & 511u, sb = (op >> 23) & 511u;
    if (kind == SV_NOP) return;
    const F2<C> a = sf.load(sa);
    F2<C> r;
    switch (kind) {
    case SV_MUL:
    case SV_SQR: {   // a square runs through the one multiplication body: a second 14 KB body costs more in instruction
                     // fetch than the product it saves
        const F2<C> b = sf.load(kind == SV_MUL ? sb : sa);
        r = sat_fp2_mul<C>(a, b);
        break;
    }
    case SV_ADD: {
        const F2<C> b = sf.load(sb);
        r.c0 = mp_add_f<C>(a.c0, b.c0);
        r.c1 = mp_add_f<C>(a.c1, b.c1);
        break;
    }
    case SV_SUB: {
        const F2<C> b = sf.load(sb);
        r.c0 = mp_sub_f<C>(a.c0, b.c0);
        r.c1 = mp_sub_f<C>(a.c1, b.c1);
        break;
    }
    case SV_XI:
        r = sat_fp2_mul_xi<C>(a);
        break;
    case SV_HALF:
        mp_half<C>(r.c0.v, a.c0.v);
        mp_half<C>(r.c1.v, a.c1.v);
        break;
    case SV_CONJ:
        r.c0 = a.c0;
        mp_neg<C>(r.c1.v, a.c1.v);
        break;
    case SV_NEG:
        mp_neg<C>(r.c0.v, a.c0.v);
        mp_neg<C>(r.c1.v, a.c1.v);
        break;
    case SV_SEL0:
    case SV_SEL1:   // the line of a pair with a point at infinity is replaced by 1 (sb = index of the pair in its group)
        r = (sf.flags >> sb) & 1u ? sf.load(SV_CONST0 + (kind == SV_SEL1 ? 1 : 0)) : a;
        break;
    default:   // SV_COPY
        r = a;
        break;
    }
    sf.store(d, r);
}

// big-endian field element -> canonical Montgomery limbs (value must be < p: the boundary's contract)
template <class C> HD LN<C::N> sv_fp_from_be(const uint8_t* be) {
    constexpr int N = C::N;
    LN<N> t, r2;
#pragma unroll
    for (int i = 0; i < N; i++) {
        const uint8_t* q = be + 4 * (N - 1 - i);
        t.v[i] = ((uint32_t)q[0] << 24) | ((uint32_t)q[1] << 16) | ((uint32_t)q[2] << 8) | q[3];
        r2.v[i] = C::R2()[i];
    }
    return sat_fp_mul<C>(t, r2);
}
template <class C> HD void sv_fp_to_be(uint8_t* be, const LN<C::N>& a) {
    constexpr int N = C::N;
    LN<N> one;
#pragma unroll
    for (int i = 0; i < N; i++) one.v[i] = i == 0 ? 1u : 0u;
    const LN<N> t = sat_fp_mul<C>(a, one);
#pragma unroll
    for (int i = 0; i < N; i++) {
        uint8_t* q = be + 4 * (N - 1 - i);
        q[0] = (uint8_t)(t.v[i] >> 24); q[1] = (uint8_t)(t.v[i] >> 16); q[2] = (uint8_t)(t.v[i] >> 8); q[3] = (uint8_t)t.v[i];
    }
}
template <class C> HD LN<C::N> sv_zero() {
    LN<C::N> z;
#pragma unroll
    for (int i = 0; i < C::N; i++) z.v[i] = 0;
    return z;
}
// stores coordinate c (0: xP, 1: yP, 2: xQ.im, 3: xQ.re, 4: yQ.im, 5: yQ.re) of pair j of the group into the state slots
// (the per-pair state of pair j follows that of pair j - 1: TX TY TZ QX QY PX PY)
template <class C, class T, int NPB> HD void sv_store_coord(const SlotFile<C, NPB>& sf, int j, int c, const LN<C::N>& v) {
    const int o = 7 * j;
    switch (c) {
    case 0: sf.store_fp(T::S_PX0 + o, 0, v); sf.store_fp(T::S_PX0 + o, 1, sv_zero<C>()); break;
    case 1: sf.store_fp(T::S_PY0 + o, 0, v); sf.store_fp(T::S_PY0 + o, 1, sv_zero<C>()); break;
    case 2: sf.store_fp(T::S_QX0 + o, 1, v); sf.store_fp(T::S_TX0 + o, 1, v); break;
    case 3: sf.store_fp(T::S_QX0 + o, 0, v); sf.store_fp(T::S_TX0 + o, 0, v); break;
    case 4: sf.store_fp(T::S_QY0 + o, 1, v); sf.store_fp(T::S_TY0 + o, 1, v); break;
    default: sf.store_fp(T::S_QY0 + o, 0, v); sf.store_fp(T::S_TY0 + o, 0, v); break;
    }
}
// f <- 1
template <class C, class T, int NPB> HD void sv_set_one(const SlotFile<C, NPB>& sf) {
    sf.store(T::S_F00, sf.load(SV_CONST0 + 1));
    const F2<C> z = sf.load(SV_CONST0);
    sf.store(T::S_F01, z); sf.store(T::S_F02, z); sf.store(T::S_F10, z); sf.store(T::S_F11, z); sf.store(T::S_F12, z);
}
// GT wire position i (w-powers 5,3,1,4,2,0) -> state slot
template <class T> HD int sv_wire_slot(int i) {
    return i == 0 ? T::S_F12 : i == 1 ? T::S_F11 : i == 2 ? T::S_F10 : i == 3 ? T::S_F02 : i == 4 ? T::S_F01 : T::S_F00;
}

// shared-memory footprint (bytes) of one block: constants, slot file, per-group flags
template <class C, class T, int NPB> constexpr size_t sv_smem_bytes() {   // NPB: groups per block
    return (size_t)T::NCONST * (2 * C::N / 4) * 16 + (size_t)T::NSLOT * (2 * C::N / 4) * NPB * 16;
}

// scratch of the in-launch product tree over nb block values with fan-in `fan`: words of all levels / number of tickets
inline size_t sv_tree_words(size_t nb, size_t fan, size_t n_limbs) {
    size_t w = 0;
    for (size_t c = nb; c > 1; c = (c + fan - 1) / fan) w += c * 12 * n_limbs;
    return w;
}
inline size_t sv_tree_counters(size_t nb, size_t fan) {
    size_t k = 0;
    for (size_t c = nb; c > 1; c = (c + fan - 1) / fan) k += (c + fan - 1) / fan;
    return k;
}

#if defined(__CUDACC__)
struct SvTables {          // device copies of one table set (code, program offsets, sequence, constants)
    const uint32_t* code;
    const uint32_t* offs;
    const uint8_t* seq;
    const uint32_t* consts;
    const uint32_t* mach_r;   // 2^(28 L) mod p (L: limbs of the machine's 28-bit form), plain N 32-bit limbs
};

// runs the code words [lo, hi) : one operation per lane and round; the only instance of the interpreter in a kernel
template <class C, int NPB, int G>
__device__ __noinline__ void sv_run(SvU4* slots, const SvU4* consts, int q, uint32_t flags, const uint32_t* code, uint32_t lo, uint32_t hi,
                                    int gl, bool act) {
    const SlotFile<C, NPB> sf{slots, consts, q, flags};
    for (uint32_t w = lo; w < hi; w += G) {
        if (act) sv_exec<C, NPB>(sf, __ldg(code + w + gl));
        if (G > 1) __syncwarp();
    }
}

// product tree over the accumulators of the block's first `ngroups` groups: f_q <- f_q * f_{q + st}; group 0 ends up with
// the product.  Every thread of the block must call it.
template <class C, class T, int NPB>
__device__ __forceinline__ void sv_block_tree(SvU4* slots, const SvU4* consts, const SvTables& tb, int q, int gl, int ngroups) {
    constexpr int G = T::G;
    const SlotFile<C, NPB> sf{slots, consts, q, 0};
    const uint32_t mlo = tb.offs[T::P_MUL12], mhi = tb.offs[T::P_MUL12 + 1];
#pragma unroll 1
    for (int st = 1; st < NPB; st <<= 1) {
        __syncthreads();
        if (st >= ngroups) break;     // uniform over the block
        const bool act = (q & (2 * st - 1)) == 0 && q + st < ngroups;
        if (act) {
            const SlotFile<C, NPB> pf{slots, consts, q + st, 0};
            for (int k = gl; k < 6; k += G) sf.store(T::S_G00 + k, pf.load(T::S_F00 + k));
        }
        if (G > 1) __syncwarp();
        sv_run<C, NPB, G>(slots, consts, q, 0, tb.code, mlo, mhi, gl, act);
    }
    __syncthreads();
}
// Internal form of an Fp12 value between the slot engine's kernels: the 6 accumulator slots F00..F12 as they are
// (canonical Montgomery limbs, R = 2^(32 N)), 12 N words per value.
// mach_l > 0: the block's value leaves in the form the machine's final-exponentiation kernel reads instead
// (machine_kernels.cuh: index 2 k + part for the coefficient of w^k, mach_l limbs of 28 bits of x * 2^(28 mach_l) mod p).
template <class C, class T, int NPB>
__device__ __forceinline__ void sv_emit_value(SvU4* slots, const SvU4* consts, const SvTables& tb, uint32_t* out, int mach_l) {
    constexpr int N = C::N;
    if (threadIdx.x >= 12) return;
    const SlotFile<C, NPB> p0{slots, consts, 0, 0};
    if (mach_l == 0) {
        const int slot = T::S_F00 + (threadIdx.x >> 1), half = threadIdx.x & 1;
        const LN<N> v = p0.load_fp(slot, half);
#pragma unroll
        for (int i = 0; i < N; i++) out[threadIdx.x * N + i] = v.v[i];
        return;
    }
    const int k = threadIdx.x >> 1, part = threadIdx.x & 1;                       // coefficient of w^k, 0 = re, 1 = im
    const int slot = (k & 1) ? T::S_F10 + (k >> 1) : T::S_F00 + (k >> 1);       // w^k: k even -> F0{k/2}, k odd -> F1{(k-1)/2}
    LN<N> r;
#pragma unroll
    for (int i = 0; i < N; i++) r.v[i] = tb.mach_r[i];
    const LN<N> x = sat_fp_mul<C>(p0.load_fp(slot, part), r);                     // x * 2^(28 L) mod p, canonical
    uint32_t* o = out + threadIdx.x * mach_l;
    for (int i = 0; i < mach_l; i++) {
        const int bit = 28 * i, w = bit >> 5, sh = bit & 31;
        const uint32_t lo = x.v[w], hi = w + 1 < N ? x.v[w + 1] : 0u;
        o[i] = (sh ? ((lo >> sh) | (hi << (32 - sh))) : lo) & 0x0FFFFFFFu;
    }
}

// Miller loops of the pairs [blockIdx.x * NPB * K, ...) and the product of ALL the launch's pairs: one Fp12 value (raw Miller
// product, no final exponentiation) in `mach_out` (machine form when mach_l > 0, else internal form).  `partials`: scratch
// for the tree levels (sv_tree_words), `counters`: zero-initialised tickets (sv_tree_counters), left zero.  WPB warps per block, G lanes per group of K pairs (one shared Miller
// accumulator), NPB = WPB * 32 / G groups per block.
// FIN: what the block that ends up with the product does with it -- SvNoFinish: nothing more (the value is in `mach_out`);
// MachFinisher<F> (machine_kernels.cuh): the final exponentiation, or the plain export, in the same launch.
struct SvNoFinish {
    struct Args { int unused; };
    static constexpr size_t SMEM_BYTES = 0;
    static constexpr bool HAS_TAIL = false;
    __device__ __forceinline__ static void run(uint32_t*, const Args&, const uint32_t*) {}
    __device__ __forceinline__ static Args for_product(const Args& a, uint8_t*, size_t) { return a; }
    __device__ __forceinline__ static const uint8_t* wire_bytes(const Args&) { return nullptr; }
    __device__ __forceinline__ static void force_false(const Args&) {}
};
// SvWireFinisher: the raw Miller product leaves as a GT wire record (12 big-endian field elements, w-powers 5,3,1,4,2,0,
// im before re) -- what a rank sends to its peers.  No shared memory beyond the slot file.
template <class C, class T> struct SvWireFinisher {
    struct Args { uint8_t* out; };
    static constexpr size_t SMEM_BYTES = 0;
    static constexpr bool HAS_TAIL = true;
    __device__ __forceinline__ static void run(uint32_t*, const Args& a, const uint32_t* in) {   // in: internal form, 12 N words
        const int t = threadIdx.x;
        if (t >= 12) return;
        const int fi = 2 * (sv_wire_slot<T>(t >> 1) - T::S_F00) + ((t & 1) ? 0 : 1);
        LN<C::N> v;
#pragma unroll
        for (int i = 0; i < C::N; i++) v.v[i] = in[fi * C::N + i];
        sv_fp_to_be<C>(a.out + (size_t)t * C::FP_BYTES, v);
    }
    __device__ __forceinline__ static Args for_product(const Args& a, uint8_t*, size_t) { return a; }
    __device__ __forceinline__ static const uint8_t* wire_bytes(const Args& a) { return a.out; }
    __device__ __forceinline__ static void force_false(const Args&) {}
};
// Multi-GPU exchange fused into the launch (bgls_miller_product_exchange_dev): the block that ends up with the product
// stores its wire-form value into a mailbox on every peer over NVLink peer memory, fences, and raises the epoch flags --
// the per-GPU partial leaves for the other ranks from inside the kernel that computed it.  world == 0: no exchange.
struct SvPeers {
    uint8_t* slot[8];
    unsigned long long* flag[8];
    int world, nbytes;
    unsigned long long epoch;
};
// Batch of independent products in one launch (throughput mode, BASELINE config 5): product c covers the pairs
// [off[c], off[c+1]).  k_slot_plan lays the blocks, tree levels and tickets of every product out (prefix sums); block b of
// k_slot_miller then finds its product by binary search.  batch.nbatch == 0: one product over all n pairs.
struct SvBatch {
    const unsigned long long* off;   // nbatch + 1 pair offsets (device)
    const unsigned* bstart;          // nbatch + 1: first block of every product
    const unsigned* lstart;          // nbatch + 1: first tree-level value of every product
    const unsigned* tstart;          // nbatch + 1: first ticket of every product
    size_t nbatch;
    uint8_t* flags8;                 // nbatch verdicts (1 = the product is the identity)
};
__device__ __forceinline__ unsigned sv_tree_values_dev(unsigned nb, unsigned fan) {
    unsigned w = 0;
    for (unsigned c = nb; c > 1; c = (c + fan - 1) / fan) w += c;
    return w;
}
__device__ __forceinline__ unsigned sv_tree_tickets_dev(unsigned nb, unsigned fan) {
    unsigned k = 0;
    for (unsigned c = nb; c > 1; c = (c + fan - 1) / fan) k += (c + fan - 1) / fan;
    return k;
}
// one block: exclusive prefix sums over the products (chunked over the threads, then combined)
__global__ void __launch_bounds__(1024) k_slot_plan(const unsigned long long* __restrict__ off, size_t nbatch, unsigned pairs_per_block, unsigned fan,
                                                    unsigned* __restrict__ bstart, unsigned* __restrict__ lstart, unsigned* __restrict__ tstart) {
    __shared__ unsigned sb[1024], sl[1024], st[1024];
    const size_t chunk = (nbatch + blockDim.x - 1) / blockDim.x;
    const size_t lo = (size_t)threadIdx.x * chunk, hi = lo + chunk < nbatch ? lo + chunk : nbatch;
    unsigned b = 0, l = 0, t = 0;
    for (size_t c = lo; c < hi; c++) {
        const unsigned nb = (unsigned)((off[c + 1] - off[c] + pairs_per_block - 1) / pairs_per_block);
        const unsigned nbe = nb ? nb : 1;   // an empty product still gets one block (it writes the verdict of the empty product)
        b += nbe; l += sv_tree_values_dev(nbe, fan); t += sv_tree_tickets_dev(nbe, fan);
    }
    sb[threadIdx.x] = b; sl[threadIdx.x] = l; st[threadIdx.x] = t;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned ab = 0, al = 0, at = 0;
        for (unsigned i = 0; i < blockDim.x; i++) {
            const unsigned xb = sb[i], xl = sl[i], xt = st[i];
            sb[i] = ab; sl[i] = al; st[i] = at;
            ab += xb; al += xl; at += xt;
        }
        bstart[nbatch] = ab; lstart[nbatch] = al; tstart[nbatch] = at;
    }
    __syncthreads();
    b = sb[threadIdx.x]; l = sl[threadIdx.x]; t = st[threadIdx.x];
    for (size_t c = lo; c < hi; c++) {
        bstart[c] = b; lstart[c] = l; tstart[c] = t;
        const unsigned nb = (unsigned)((off[c + 1] - off[c] + pairs_per_block - 1) / pairs_per_block);
        const unsigned nbe = nb ? nb : 1;
        b += nbe; l += sv_tree_values_dev(nbe, fan); t += sv_tree_tickets_dev(nbe, fan);
    }
}

template <class C, class T, int WPB, class FIN>
__global__ void __launch_bounds__(WPB * 32) k_slot_miller(SvTables tb, const uint8_t* __restrict__ g1, const uint8_t* __restrict__ g2,
                                                          size_t n, uint32_t* __restrict__ partials, unsigned* __restrict__ counters,
                                                          uint32_t* __restrict__ mach_out, int mach_l, typename FIN::Args fin,
                                                          SvBatch batch, unsigned long long* __restrict__ trace, SvPeers peers) {
    constexpr int N = C::N, FB = C::FP_BYTES, G = T::G, K = T::K, NPB = WPB * 32 / G, W4 = 2 * N / 4;
    extern __shared__ uint4 sv_sm[];
    // ---- which product, which block of it
    size_t blk = blockIdx.x, nblk = gridDim.x, pair0 = 0;
    if (batch.nbatch) {
        if (blockIdx.x >= batch.bstart[batch.nbatch]) return;
        size_t lo = 0, hi = batch.nbatch;            // last c with bstart[c] <= blockIdx.x
        while (hi - lo > 1) {
            const size_t mid = (lo + hi) / 2;
            if (batch.bstart[mid] <= blockIdx.x) lo = mid; else hi = mid;
        }
        blk = blockIdx.x - batch.bstart[lo];
        nblk = batch.bstart[lo + 1] - batch.bstart[lo];
        pair0 = batch.off[lo];
        n = batch.off[lo + 1] - pair0;
        g1 += pair0 * 2 * FB;
        g2 += pair0 * 4 * FB;
        partials += (size_t)batch.lstart[lo] * 12 * N;
        counters += batch.tstart[lo];
        mach_out += lo * 12 * (mach_l ? mach_l : N);
        fin = FIN::for_product(fin, batch.flags8, lo);
    }
    unsigned long long t_start = 0;
    if (trace && threadIdx.x == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_start));
    // the program words stay in global memory: a few KB read by every warp of the GPU, i.e. L1 hits; staging them per
    // block cost 3 KB of shared memory (a fifth of a block's footprint at 8 lanes per group)
    const uint32_t* code = tb.code;
    SvU4* consts = (SvU4*)sv_sm;
    SvU4* slots = consts + T::NCONST * W4;
    for (int i = threadIdx.x; i < T::NCONST * 2 * N; i += blockDim.x) ((uint32_t*)consts)[i] = tb.consts[i];
    // lane -> (group, lane of the group): the 32 / G groups of a warp are interleaved, so that the lanes of a quarter
    // warp hold the SAME lane index of consecutive groups: they run the same operation on the same slot and their
    // 16-byte accesses are consecutive (conflict free; with lane = group * G + index every access was a G-way conflict)
    constexpr int GPW = 32 / G;
    const int lane = threadIdx.x & 31;
    const int q = (threadIdx.x >> 5) * GPW + lane % GPW, gl = lane / GPW;
    const size_t first = blk * NPB * K;                                 // first pair of the block (relative to its product)
    const size_t ngroups_left = n > first ? (n - first + K - 1) / K : 0;
    const int ngroups = (int)(ngroups_left < (size_t)NPB ? ngroups_left : (size_t)NPB);
    SlotFile<C, NPB> sf{slots, consts, q, 0};
    // ---- inputs: G1 = x || y, G2 = x_im || x_re || y_im || y_re (big-endian); coordinate c of pair j is converted by lane
    // (6 j + c) % G.  A point is infinity when its record is all zero (or carries the bls12 infinity flag).
    uint32_t anyp = 0, anyq = 0, flag = 0;   // bit j: pair j has a non-zero G1 / G2 coordinate / an infinity flag
#pragma unroll 1
    for (int t = gl; t < 6 * K; t += G) {
        const int j = t / 6, c = t % 6;
        const size_t pair = first + (size_t)q * K + j;
        LN<N> v = sv_zero<C>();
        if (pair < n) {
            const uint8_t* src = c < 2 ? g1 + pair * 2 * FB + c * FB : g2 + pair * 4 * FB + (c - 2) * FB;
            v = sv_fp_from_be<C>(src);
            uint32_t any = 0;
#pragma unroll
            for (int i = 0; i < N; i++) any |= v.v[i];
            if (any) { if (c < 2) anyp |= 1u << j; else anyq |= 1u << j; }
            if (!C::IS_BN && (c == 0 || c == 2) && (src[0] & 0x40)) flag |= 1u << j;
        }
        sv_store_coord<C, T, NPB>(sf, j, c, v);
    }
#pragma unroll
    for (int o = GPW; o < 32; o <<= 1) {   // over the lanes of the group
        anyp |= __shfl_xor_sync(0xFFFFFFFFu, anyp, o);
        anyq |= __shfl_xor_sync(0xFFFFFFFFu, anyq, o);
        flag |= __shfl_xor_sync(0xFFFFFFFFu, flag, o);
    }
    const uint32_t inf = (~(anyp & anyq) | flag) & ((1u << K) - 1);     // pairs beyond n have all-zero coordinates
    sf.flags = inf;
    __syncthreads();   // constants staged, inputs stored
    for (int j = gl; j < K; j += G) sf.store(T::S_TZ0 + 7 * j, sf.load(SV_CONST0 + 1));
    __syncwarp();
    // ---- the Miller loop: a sequence of programs
#pragma unroll 1
    for (int s = 0; s < T::SEQ_LEN; s++) {
        const uint32_t pid = tb.seq[s];
        sv_run<C, NPB, G>(slots, consts, q, inf, code, tb.offs[pid], tb.offs[pid + 1], gl, true);
    }
    if (K == 1 && inf && gl == 0) sv_set_one<C, T, NPB>(sf);            // K > 1: the lines of such pairs were replaced by 1
    // (a block without pairs -- the empty product -- still ran the programs on all-infinity groups: group 0 holds 1)
    sv_block_tree<C, T, NPB>(slots, consts, tb, q, gl, ngroups);
    if (trace && threadIdx.x == 0) {   // BGLS_TRACE: (start ns, end ns, SM) of every Miller block, for timeline reconstruction
        unsigned long long t_end;
        unsigned smid;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_end));
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        const unsigned long long slot = atomicAdd(trace, 1ull);
        if (slot < (1ull << 20)) { trace[1 + 3 * slot] = t_start; trace[2 + 3 * slot] = t_end; trace[3 + 3 * slot] = smid; }
    }
    // ---- cross-block product tree inside the same launch: a block stores its value, and the LAST block of every group of
    // FAN blocks to arrive (atomic ticket) multiplies the group's values and moves one level up; the block that ends up
    // with the last value writes it in the machine's form.  No dependent kernel is needed for the tree: with many
    // products in flight every dependent launch in a stream stalls the hardware queue it shares with other streams
    // (measured: 2.6 M pairings/s with three tree launches per product, 4.1 M without them).
    constexpr int FAN = 2 * NPB;
    __shared__ int s_last;
    size_t count = nblk, idx = blk;
    uint32_t* lvl = partials;
    unsigned* cnt = counters;
#pragma unroll 1
    while (count > 1) {
        sv_emit_value<C, T, NPB>(slots, consts, tb, lvl + idx * 12 * N, 0);
        __threadfence();
        __syncthreads();
        const size_t grp = idx / FAN;
        const int size = (int)(count - grp * FAN < (size_t)FAN ? count - grp * FAN : (size_t)FAN);
        if (threadIdx.x == 0) {
            const unsigned t = atomicAdd(cnt + grp, 1u);
            s_last = t == (unsigned)(size - 1);
            if (s_last) cnt[grp] = 0;      // ready for the next launch
        }
        __syncthreads();
        if (!s_last) return;
        __threadfence();
        // group q: F <- value 2q, G <- value 2q + 1 of this group of blocks
        for (int t = gl; t < 24; t += G) {
            const int which = t / 12, fi = t % 12, v = 2 * q + which;
            if (v < size) {
                LN<N> x;
                const uint32_t* src = lvl + (grp * FAN + v) * 12 * N + fi * N;
#pragma unroll
                for (int i = 0; i < N; i++) x.v[i] = __ldcg(src + i);
                sf.store_fp((which ? T::S_G00 : T::S_F00) + (fi >> 1), fi & 1, x);
            }
        }
        __syncthreads();
        sv_run<C, NPB, G>(slots, consts, q, 0, code, tb.offs[T::P_MUL12], tb.offs[T::P_MUL12 + 1], gl, 2 * q + 1 < size);
        sv_block_tree<C, T, NPB>(slots, consts, tb, q, gl, (size + 1) / 2);
        const size_t groups = (count + FAN - 1) / FAN;
        lvl += count * 12 * N;
        cnt += groups;
        idx = grp;
        count = groups;
    }
    sv_emit_value<C, T, NPB>(slots, consts, tb, mach_out, mach_l);
    if (FIN::HAS_TAIL) {   // the final exponentiation (or the export) runs on the first warp of this block, over the same shared memory
        __threadfence_block();
        __syncthreads();
        if (threadIdx.x < 32) FIN::run((uint32_t*)sv_sm, fin, mach_out);
        if (peers.world > 0) {   // FIN exported the wire bytes: payload to every peer, then the flags
            __threadfence();
            __syncthreads();
            const uint8_t* src = FIN::wire_bytes(fin);
            for (int idx = threadIdx.x; idx < peers.world * peers.nbytes; idx += blockDim.x) {
                const int r = idx / peers.nbytes, b = idx % peers.nbytes;
                peers.slot[r][b] = src[b];
            }
            __threadfence_system();
            __syncthreads();
            if ((int)threadIdx.x < peers.world) {
                *(volatile unsigned long long*)peers.flag[threadIdx.x] = peers.epoch;
                __threadfence_system();
            }
        }
    }
}

// Finishing step of a sharded verification in ONE launch: the product of k <= 2 NPB Miller values in wire form (the
// all-gathered partials of the ranks, /root/reference/curves/curve.go:141-169 across GPUs) and FIN on it -- the final
// exponentiation and the comparison with the identity, or the plain export.  Replaces import + tree + finish (three
// dependent launches per step, which cost the multi-GPU pipeline a tenth of its throughput).
template <class C, class T, int WPB, class FIN>
__global__ void __launch_bounds__(WPB * 32) k_slot_finish_bytes(SvTables tb, const uint8_t* vals, int k, uint32_t* __restrict__ mach_out,
                                                                int mach_l, typename FIN::Args fin, const unsigned long long* wait_flags,
                                                                unsigned long long epoch, int* err) {
    constexpr int N = C::N, FB = C::FP_BYTES, G = T::G, NPB = WPB * 32 / G, W4 = 2 * N / 4, GPW = 32 / G;
    extern __shared__ uint4 sv_sm[];
    SvU4* consts = (SvU4*)sv_sm;
    SvU4* slots = consts + T::NCONST * W4;
    for (int i = threadIdx.x; i < T::NCONST * 2 * N; i += blockDim.x) ((uint32_t*)consts)[i] = tb.consts[i];
    const int lane = threadIdx.x & 31;
    const int q = (threadIdx.x >> 5) * GPW + lane % GPW, gl = lane / GPW;
    const SlotFile<C, NPB> sf{slots, consts, q, 0};
    if (wait_flags) {   // peer-memory exchange: the k values are the mailbox records of the ranks; wait (bounded) for their flags
        if ((int)threadIdx.x < k) {
            const volatile unsigned long long* f = wait_flags + threadIdx.x;
            long long spins = 0;
            while (*f < epoch) {
                if (++spins > (1ll << 27)) { *err = 1; break; }   // several seconds: a missing peer is an error, not a hung GPU
                __nanosleep(64);
            }
        }
        __threadfence_system();
    }
    __syncthreads();
    // group q: F <- value 2q, G <- value 2q + 1.  Wire position t (coefficient t / 2 of sv_wire_slot, im before re).
    for (int t = gl; t < 24; t += G) {
        const int which = t / 12, w = t % 12, v = 2 * q + which;
        if (v < k) {
            uint8_t be[FB];
            const volatile uint8_t* src = vals + ((size_t)v * 12 + w) * FB;   // written by peers: not through the read-only path
            for (int i = 0; i < FB; i++) be[i] = src[i];
            const LN<N> x = sv_fp_from_be<C>(be);
            sf.store_fp(sv_wire_slot<T>(w >> 1) + (which ? T::S_G00 - T::S_F00 : 0), (w & 1) ? 0 : 1, x);
        }
    }
    if (k == 0 && q == 0 && gl == 0) sv_set_one<C, T, NPB>(sf);   // the empty product
    __syncthreads();
    sv_run<C, NPB, G>(slots, consts, q, 0, tb.code, tb.offs[T::P_MUL12], tb.offs[T::P_MUL12 + 1], gl, 2 * q + 1 < k);
    sv_block_tree<C, T, NPB>(slots, consts, tb, q, gl, (k + 1) / 2);
    sv_emit_value<C, T, NPB>(slots, consts, tb, mach_out, mach_l);
    if (FIN::HAS_TAIL) {
        __threadfence_block();
        __syncthreads();
        if (threadIdx.x < 32) FIN::run((uint32_t*)sv_sm, fin, mach_out);
        if (wait_flags) {   // a wait that timed out: the verdict is false whatever the stale records multiplied to
            __syncthreads();
            if (threadIdx.x == 0 && *err) FIN::force_false(fin);
        }
    }
}

#endif

}  // namespace bgls
