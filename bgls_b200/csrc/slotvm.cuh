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

enum : uint32_t { SV_NOP = 0, SV_MUL, SV_SQR, SV_ADD, SV_SUB, SV_XI, SV_HALF, SV_CONJ, SV_NEG, SV_COPY };
constexpr uint32_t SV_CONST0 = 256;

struct alignas(16) SvU4 { uint32_t x, y, z, w; };   // 16-byte unit of the slot file (uint4 on the device)

template <class C, int NPB> struct SlotFile {
    static constexpr int N = C::N, W4 = 2 * N / 4;
    SvU4* slots;          // pair slots
    const SvU4* consts;   // constants, [c][W4]
    int q;                // pair index inside the block

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
// stores coordinate c (0: xP, 1: yP, 2: xQ.im, 3: xQ.re, 4: yQ.im, 5: yQ.re) of a pair into the state slots
template <class C, class T, int NPB> HD void sv_store_coord(const SlotFile<C, NPB>& sf, int c, const LN<C::N>& v) {
    switch (c) {
    case 0: sf.store_fp(T::S_PX, 0, v); sf.store_fp(T::S_PX, 1, sv_zero<C>()); break;
    case 1: sf.store_fp(T::S_PY, 0, v); sf.store_fp(T::S_PY, 1, sv_zero<C>()); break;
    case 2: sf.store_fp(T::S_QX, 1, v); sf.store_fp(T::S_TX, 1, v); break;
    case 3: sf.store_fp(T::S_QX, 0, v); sf.store_fp(T::S_TX, 0, v); break;
    case 4: sf.store_fp(T::S_QY, 1, v); sf.store_fp(T::S_TY, 1, v); break;
    default: sf.store_fp(T::S_QY, 0, v); sf.store_fp(T::S_TY, 0, v); break;
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

// shared-memory footprint (bytes) of one block: code words, constants, slot file, per-pair flags
template <class C, class T, int NPB> constexpr size_t sv_smem_bytes() {
    return ((size_t)T::NWORDS * 4 + 15) / 16 * 16 + (size_t)T::NCONST * (2 * C::N / 4) * 16 + (size_t)T::NSLOT * (2 * C::N / 4) * NPB * 16;
}

#if defined(__CUDACC__)
struct SvTables {          // device copies of one table set (code, program offsets, sequence, constants)
    const uint32_t* code;
    const uint32_t* offs;
    const uint8_t* seq;
    const uint32_t* consts;
};

// runs the code words [lo, hi) : one operation per lane and round; the only instance of the interpreter in a kernel
template <class C, int NPB, int G>
__device__ __noinline__ void sv_run(SvU4* slots, const SvU4* consts, int q, const uint32_t* code, uint32_t lo, uint32_t hi, int gl, bool act) {
    const SlotFile<C, NPB> sf{slots, consts, q};
    for (uint32_t w = lo; w < hi; w += G) {
        if (act) sv_exec<C, NPB>(sf, code[w + gl]);
        if (G > 1) __syncwarp();
    }
}

// Miller loops of the pairs [blockIdx.x * NPB, ...) and their product: one GT wire record (raw Miller product, no
// final exponentiation) per block.  WPB warps per block, G lanes per pair, NPB = WPB * 32 / G pairs per block.
template <class C, class T, int WPB>
__global__ void __launch_bounds__(WPB * 32) k_slot_miller(SvTables tb, const uint8_t* __restrict__ g1, const uint8_t* __restrict__ g2,
                                                          size_t n, uint8_t* __restrict__ partials) {
    constexpr int N = C::N, FB = C::FP_BYTES, G = T::G, NPB = WPB * 32 / G, W4 = 2 * N / 4;
    extern __shared__ uint4 sv_sm[];
    uint32_t* code = (uint32_t*)sv_sm;
    SvU4* consts = (SvU4*)(sv_sm + (T::NWORDS * 4 + 15) / 16);
    SvU4* slots = consts + T::NCONST * W4;
    for (int i = threadIdx.x; i < T::NWORDS; i += blockDim.x) code[i] = tb.code[i];
    for (int i = threadIdx.x; i < T::NCONST * 2 * N; i += blockDim.x) ((uint32_t*)consts)[i] = tb.consts[i];
    const int q = threadIdx.x / G, gl = threadIdx.x % G;
    const size_t first = (size_t)blockIdx.x * NPB;
    const size_t pair = first + q;
    const int npairs = (int)(n - first < (size_t)NPB ? n - first : (size_t)NPB);
    const SlotFile<C, NPB> sf{slots, consts, q};
    // ---- inputs: G1 = x || y, G2 = x_im || x_re || y_im || y_re (big-endian); coordinate c is converted by lane c % G
    uint32_t anyp = 0, anyq = 0, flag = 0;
    if (pair < n) {
        const uint8_t* r1 = g1 + pair * 2 * FB;
        const uint8_t* r2 = g2 + pair * 4 * FB;
#pragma unroll 1
        for (int c = gl; c < 6; c += G) {
            const uint8_t* src = c < 2 ? r1 + c * FB : r2 + (c - 2) * FB;
            const LN<N> v = sv_fp_from_be<C>(src);
            uint32_t any = 0;
#pragma unroll
            for (int i = 0; i < N; i++) any |= v.v[i];
            if (c < 2) anyp |= any; else anyq |= any;
            if (!C::IS_BN && (c == 0 || c == 2) && (src[0] & 0x40)) flag = 1;
            sv_store_coord<C, T, NPB>(sf, c, v);
        }
    }
    // a point is infinity when its record is all zero (or carries the bls12 infinity flag): the pair contributes 1
#pragma unroll
    for (int o = 1; o < G; o <<= 1) {
        anyp |= __shfl_xor_sync(0xFFFFFFFFu, anyp, o);
        anyq |= __shfl_xor_sync(0xFFFFFFFFu, anyq, o);
        flag |= __shfl_xor_sync(0xFFFFFFFFu, flag, o);
    }
    const bool inf = pair >= n || anyp == 0 || anyq == 0 || flag != 0;
    __syncthreads();   // code and constants staged, inputs stored
    if (gl == 0) sf.store(T::S_TZ, sf.load(SV_CONST0 + 1));
    __syncwarp();
    // ---- the Miller loop: a sequence of programs
#pragma unroll 1
    for (int s = 0; s < T::SEQ_LEN; s++) {
        const uint32_t pid = tb.seq[s];
        sv_run<C, NPB, G>(slots, consts, q, code, tb.offs[pid], tb.offs[pid + 1], gl, true);
    }
    if (inf && gl == 0) sv_set_one<C, T, NPB>(sf);
    // ---- product tree over the block's pairs: f_q <- f_q * f_{q + st}
    const uint32_t mlo = tb.offs[T::P_MUL12], mhi = tb.offs[T::P_MUL12 + 1];
#pragma unroll 1
    for (int st = 1; st < NPB; st <<= 1) {
        __syncthreads();
        if (st >= npairs) break;     // uniform over the block
        const bool act = (q & (2 * st - 1)) == 0 && q + st < npairs;
        if (act) {
            const SlotFile<C, NPB> pf{slots, consts, q + st};
            for (int k = gl; k < 6; k += G) sf.store(T::S_G00 + k, pf.load(T::S_F00 + k));
        }
        if (G > 1) __syncwarp();
        sv_run<C, NPB, G>(slots, consts, q, code, mlo, mhi, gl, act);
    }
    __syncthreads();
    // ---- wire record of the block's product (pair 0): w-powers 5,3,1,4,2,0, each (im, re)
    if (threadIdx.x < 12) {
        const int i = threadIdx.x >> 1, part = threadIdx.x & 1;          // position i, part 0 = im, 1 = re
        const SlotFile<C, NPB> p0{slots, consts, 0};
        sv_fp_to_be<C>(partials + ((size_t)blockIdx.x * 12 + threadIdx.x) * FB, p0.load_fp(sv_wire_slot<T>(i), part ? 0 : 1));
    }
}
#endif

}  // namespace bgls
