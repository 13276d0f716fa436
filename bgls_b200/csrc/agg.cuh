// Point aggregation on six lanes per addition: AggregatePoints (/root/reference/curves/curve.go:73-121, the n-way
// Point.Add fan-in of verifyMultiSignature, bgls/bgls.go:89-92) as ONE launch.
//
// n points are too few for one-addition-per-thread code to fill a B200 (65,536 keys against 75,776 resident threads,
// and a thread-serial G2 addition is a 20 us dependent chain), so one addition is spread over the six lanes of a
// group: the complete projective addition for y^2 = x^3 + b (Renes-Costello-Batina 2016, algorithm 7; exception
// free on these odd-order groups, so infinity, doubling and P + (-P) need no branches) is two layers of six
// independent field multiplications with a short linear layer between them,
//
//   layer 1   t0 = X1 X2   t1 = Y1 Y2   t2 = Z1 Z2   t3 = (X1+Y1)(X2+Y2)   t4 = (Y1+Z1)(Y2+Z2)   t5 = (X1+Z1)(X2+Z2)
//   linear    A = t3-t0-t1   B = t4-t1-t2   Cb = 3b (t5-t0-t2)   D = 3 t0   E = t1 + 3b t2   F = t1 - 3b t2
//   layer 2   X3 = A F - B Cb      Y3 = F E + Cb D      Z3 = E B + D A
//
// one multiplication per lane and layer, operands exchanged through a per-group slot file in shared memory (same
// interleaved layout as slotvm.cuh).  Field arithmetic is sat.cuh's (saturated 32-bit limbs, canonical values).
//
// Inputs are NOT converted to Montgomery form: a projective point is scale invariant and every output coordinate is
// bi-homogeneous of degree (2, 2) in the two inputs, so plain limbs (the point scaled by R^-1) go straight into the
// formulas; only the curve constant 3b is a Montgomery constant.  The one inversion at the very end is a binary
// (Kaliski) inversion, 30 k instructions instead of a 570-multiplication Fermat power.
//
// Phases of k_agg: (1) every group adds its strided share of the points into its accumulator; (2) binary tree over
// the groups of the block; (3) cross-block tree inside the same launch (atomic tickets, the last block of every
// FAN to arrive goes on -- the scheme of k_slot_miller); (4) the last block converts to affine and writes the record.
//
// Host build (tests/host_emul): the same stage functions run lane by lane.
#pragma once
#include <cstdint>

#include "inv.cuh"
#include "sat.cuh"
#include "slotvm.cuh"

namespace bgls {

// a^-1 mod p as a plain integer, 0 < a < p: binary inversion (inv.cuh), then the division by 2^k as two Montgomery products
template <class C> HD LN<C::N> mp_inv_plain(const LN<C::N>& a) {
    constexpr int N = C::N;
    LN<N> x, e1, e2;
    const int k = mp_almost_inv<C>(x.v, a.v);
    inv_shift_limbs<N>(e1.v, e2.v, k);
    return sat_fp_mul<C>(sat_fp_mul<C>(x, e1), e2);
}

// ---------------------------------------------------------------- the two element types
template <class C> HD LN<C::N> agg_small_mul(const LN<C::N>& a, int twelve) {   // 12 a (bls12-381) or 9 a (altbn128)
    const LN<C::N> x2 = mp_add_f<C>(a, a), x4 = mp_add_f<C>(x2, x2), x8 = mp_add_f<C>(x4, x4);
    return mp_add_f<C>(x8, twelve ? x4 : a);
}
template <class C> struct AggFp {                 // G1: y^2 = x^3 + b over Fp, 3b = 9 (altbn128) / 12 (bls12-381)
    using Curve = C;
    using T = LN<C::N>;
    static constexpr int HALVES = 1;
    static HD LN<C::N>& half(T& a, int) { return a; }
    static HD const LN<C::N>& half(const T& a, int) { return a; }
    static HD T mul(const T& a, const T& b) { return mp_redc_f<C>(mp_mul_f<C::N>(a, b)); }
    static HD T add(const T& a, const T& b) { return mp_add_f<C>(a, b); }
    static HD T sub(const T& a, const T& b) { return mp_sub_f<C>(a, b); }
    static constexpr bool B3_MUL = false;
    static HD T mul_b3(const T& a) { return agg_small_mul<C>(a, !C::IS_BN); }
    static HD T b3() { return T{}; }
};
template <class C> struct AggFp2 {                // G2: the twist over Fp2
    using Curve = C;
    using T = F2<C>;
    static constexpr int HALVES = 2;
    static HD LN<C::N>& half(T& a, int h) { return h ? a.c1 : a.c0; }
    static HD const LN<C::N>& half(const T& a, int h) { return h ? a.c1 : a.c0; }
    static HD T mul(const T& a, const T& b) { return sat_fp2_mul<C>(a, b); }
    static HD T add(const T& a, const T& b) { T r; r.c0 = mp_add_f<C>(a.c0, b.c0); r.c1 = mp_add_f<C>(a.c1, b.c1); return r; }
    static HD T sub(const T& a, const T& b) { T r; r.c0 = mp_sub_f<C>(a.c0, b.c0); r.c1 = mp_sub_f<C>(a.c1, b.c1); return r; }
    static constexpr bool B3_MUL = C::IS_BN;      // altbn128: 3 b' = 9 / (9 + i), a general constant (Montgomery form)
    static HD T b3() {
        T k;
#pragma unroll
        for (int i = 0; i < C::N; i++) { k.c0.v[i] = C::B2X3()[i]; k.c1.v[i] = C::B2X3()[C::N + i]; }
        return k;
    }
    static HD T mul_b3(const T& a) {              // bls12-381: 3 b' = 12 (1 + i)
        T r;
        r.c0 = agg_small_mul<C>(mp_sub_f<C>(a.c0, a.c1), 1);
        r.c1 = agg_small_mul<C>(mp_add_f<C>(a.c0, a.c1), 1);
        return r;
    }
};

// ---------------------------------------------------------------- per-group slot file
enum : int { AG_X1 = 0, AG_Y1, AG_Z1, AG_X2, AG_Y2, AG_Z2, AG_T0, AG_M0 = AG_T0 + 6, AG_NSLOT = AG_M0 + 6 };
constexpr int AGG_LANES = 6, AGG_GPW = 5;         // lanes per addition, groups per warp (lanes 30, 31 idle)

template <class E, int NGB> struct AggFile {      // NGB: groups per block
    using C = typename E::Curve;
    static constexpr int N = C::N, W4 = E::HALVES * N / 4;
    SvU4* slots;
    int q;
    HD LN<N> load_half(int s, int h) const {
        const SvU4* p = slots + ((size_t)s * W4 + h * (N / 4)) * NGB + q;
        LN<N> r;
#pragma unroll
        for (int w = 0; w < N / 4; w++) {
            const SvU4 t = p[w * NGB];
            r.v[4 * w] = t.x; r.v[4 * w + 1] = t.y; r.v[4 * w + 2] = t.z; r.v[4 * w + 3] = t.w;
        }
        return r;
    }
    HD void store_half(int s, int h, const LN<N>& a) const {
        SvU4* p = slots + ((size_t)s * W4 + h * (N / 4)) * NGB + q;
#pragma unroll
        for (int w = 0; w < N / 4; w++) {
            SvU4 t;
            t.x = a.v[4 * w]; t.y = a.v[4 * w + 1]; t.z = a.v[4 * w + 2]; t.w = a.v[4 * w + 3];
            p[w * NGB] = t;
        }
    }
    HD typename E::T load(int s) const {
        typename E::T r;
#pragma unroll
        for (int h = 0; h < E::HALVES; h++) E::half(r, h) = load_half(s, h);
        return r;
    }
    HD void store(int s, const typename E::T& a) const {
#pragma unroll
        for (int h = 0; h < E::HALVES; h++) store_half(s, h, E::half(a, h));
    }
};
template <class E, int NGB> constexpr size_t agg_smem_bytes() {
    return (size_t)AG_NSLOT * (E::HALVES * E::Curve::N / 4) * NGB * 16;
}

// ---- one addition (X1:Y1:Z1) += (X2:Y2:Z2) as three steps, each a multiplication part and a linear part with a group
// barrier after either; lane `r` of the group.  The multiplications of all steps go through ONE call site (agg_add /
// the host emulator loop over `step`): two inlined copies of a 30 KB Fp2 multiplication body per addition, times the
// three places an addition is made from, do not fit the instruction cache, and a lone warp then waits on instruction
// fetch (measured: 11 us per tree level against 6).
//   step 0   t_r = layer-1 products                          linear: A, B, C, D -> M0..M3
//   step 1   M2 <- 3b C (lane 2), t2 <- 3b t2 (lane 4)       linear: E = t1 + t2 -> M4, F = t1 - t2 -> M5
//   step 2   layer-2 products -> t_r                         linear: X1 = t0 - t1, Y1 = t2 + t3, Z1 = t4 + t5
template <class E> HD bool agg_mul_active(int step, int r) { return step != 1 || (E::B3_MUL && (r == 2 || r == 4)); }
template <class E, class F> HD void agg_mul_operands(const F& f, int step, int r, typename E::T& a, typename E::T& b) {
    if (step == 0) {
        const int i = r < 3 ? r : (r == 4 ? 1 : 0), j = r == 3 ? 1 : 2;   // sums: r = 3: X+Y, 4: Y+Z, 5: X+Z
        a = f.load(AG_X1 + i);
        b = f.load(AG_X2 + i);
        if (r >= 3) {
            a = E::add(a, f.load(AG_X1 + j));
            b = E::add(b, f.load(AG_X2 + j));
        }
    } else if (step == 1) {
        a = f.load(r == 2 ? AG_M0 + 2 : AG_T0 + 2);
        b = E::b3();
    } else {                                      // A F, B Cb, F E, Cb D, E B, D A   (A..F = M0..M5)
        const int ia = r == 0 ? 0 : r == 1 ? 1 : r == 2 ? 5 : r == 3 ? 2 : r == 4 ? 4 : 3;
        const int ib = r == 0 ? 5 : r == 1 ? 2 : r == 2 ? 4 : r == 3 ? 3 : r == 4 ? 1 : 0;
        a = f.load(AG_M0 + ia);
        b = f.load(AG_M0 + ib);
    }
}
HD int agg_mul_dst(int step, int r) { return step == 1 ? (r == 2 ? AG_M0 + 2 : AG_T0 + 2) : AG_T0 + r; }
// step 1 on the curves whose 3b is a small multiple (of 1 or of 1 + i): additions only
template <class E, class F> HD void agg_b3_small(const F& f, int r) {
    if (r != 2 && r != 4) return;
    const int s = agg_mul_dst(1, r);
    f.store(s, E::mul_b3(f.load(s)));
}
template <class E, class F> HD void agg_linear(const F& f, int step, int r) {
    if (step == 0) {
        if (r >= 4) return;
        typename E::T w = f.load(AG_T0 + (r < 3 ? 3 + r : 0));
        if (r < 3) {                              // A, B, C: t3 - t0 - t1, t4 - t1 - t2, t5 - t0 - t2
            w = E::sub(w, f.load(AG_T0 + (r == 1 ? 1 : 0)));
            w = E::sub(w, f.load(AG_T0 + (r == 0 ? 1 : 2)));
        } else {                                  // D = 3 t0
            const typename E::T t0 = w;
            w = E::add(w, t0);
            w = E::add(w, t0);
        }
        f.store(AG_M0 + r, w);
    } else if (step == 1) {
        if (r < 4) return;
        const typename E::T t1 = f.load(AG_T0 + 1), u = f.load(AG_T0 + 2);
        f.store(AG_M0 + r, r == 4 ? E::add(t1, u) : E::sub(t1, u));
    } else {
        if (r >= 3) return;
        const typename E::T a = f.load(AG_T0 + 2 * r), b = f.load(AG_T0 + 2 * r + 1);
        f.store(AG_X1 + r, r == 0 ? E::sub(a, b) : E::add(a, b));
    }
}

// ---- inputs
// big-endian field element -> plain limbs (no Montgomery conversion, see the header)
template <class C> HD LN<C::N> agg_read_be(const uint8_t* be) {
    constexpr int N = C::N;
    LN<N> t;
#if defined(__CUDA_ARCH__)
    if (((uintptr_t)be & 15) == 0) {
#pragma unroll
        for (int w = 0; w < N / 4; w++) {
            const uint4 x = __ldg((const uint4*)be + w);
            t.v[N - 1 - 4 * w] = __byte_perm(x.x, 0, 0x0123);
            t.v[N - 2 - 4 * w] = __byte_perm(x.y, 0, 0x0123);
            t.v[N - 3 - 4 * w] = __byte_perm(x.z, 0, 0x0123);
            t.v[N - 4 - 4 * w] = __byte_perm(x.w, 0, 0x0123);
        }
        return t;
    }
#endif
#pragma unroll
    for (int i = 0; i < N; i++) {
        const uint8_t* q = be + 4 * (N - 1 - i);
        t.v[i] = ((uint32_t)q[0] << 24) | ((uint32_t)q[1] << 16) | ((uint32_t)q[2] << 8) | q[3];
    }
    return t;
}
template <class C> HD void agg_write_be(uint8_t* be, const LN<C::N>& a) {
#pragma unroll
    for (int i = 0; i < C::N; i++) {
        uint8_t* q = be + 4 * (C::N - 1 - i);
        q[0] = (uint8_t)(a.v[i] >> 24); q[1] = (uint8_t)(a.v[i] >> 16); q[2] = (uint8_t)(a.v[i] >> 8); q[3] = (uint8_t)a.v[i];
    }
}
template <int N> HD LN<N> agg_small(uint32_t x) {
    LN<N> r;
#pragma unroll
    for (int i = 0; i < N; i++) r.v[i] = i == 0 ? x : 0u;
    return r;
}
// Coordinate c of a wire record (G1: x, y; G2: x_im, x_re, y_im, y_re) goes to input slot AG_X2 + c / HALVES; lane
// `role` stores coordinate `role` (role < 2 HALVES) and lane 2 HALVES the Z coordinate.  Infinity becomes (0 : 1 : 0).
template <class E, class F> HD void agg_store_input(const F& f, int role, const LN<E::Curve::N>& v, bool inf) {
    constexpr int N = E::Curve::N, NC = 2 * E::HALVES;
    if (role < NC) {
        const int half = E::HALVES == 2 ? 1 - (role & 1) : 0;
        f.store_half(AG_X2 + role / E::HALVES, half, inf ? agg_small<N>(role == NC - 1 ? 1u : 0u) : v);
    } else if (role == NC) {
        f.store_half(AG_Z2, 0, agg_small<N>(inf ? 0u : 1u));
        if (E::HALVES == 2) f.store_half(AG_Z2, 1, agg_small<N>(0u));
    }
}
// accumulator <- (0 : 1 : 0)
template <class E, class F> HD void agg_init_acc(const F& f, int role) {
    constexpr int N = E::Curve::N;
    if (role < 3) {
        f.store_half(AG_X1 + role, 0, agg_small<N>(role == 1 ? 1u : 0u));
        if (E::HALVES == 2) f.store_half(AG_X1 + role, 1, agg_small<N>(0u));
    }
}
// projective value <-> 3 HALVES N words; lane `role` moves half `role` (role < 3 HALVES)
template <class E, class F> HD void agg_export(const F& f, int role, int first_slot, uint32_t* dst) {
    constexpr int N = E::Curve::N;
    if (role >= 3 * E::HALVES) return;
    const LN<N> v = f.load_half(first_slot + role / E::HALVES, role % E::HALVES);
#pragma unroll
    for (int i = 0; i < N; i++) dst[role * N + i] = v.v[i];
}
template <class E, class F> HD void agg_import(const F& f, int role, int first_slot, const uint32_t* src) {
    constexpr int N = E::Curve::N;
    if (role >= 3 * E::HALVES) return;
    LN<N> v;
#pragma unroll
    for (int i = 0; i < N; i++) {
#if defined(__CUDA_ARCH__)
        v.v[i] = __ldcg(src + role * N + i);
#else
        v.v[i] = src[role * N + i];
#endif
    }
    f.store_half(first_slot + role / E::HALVES, role % E::HALVES, v);
}
// the accumulator of the group as an affine wire record (thread serial: one inversion)
template <class E, class F> HD void agg_store_affine(const F& f, uint8_t* out) {
    using C = typename E::Curve;
    constexpr int N = C::N, FB = C::FP_BYTES;
    LN<N> r2;
#pragma unroll
    for (int i = 0; i < N; i++) r2.v[i] = C::R2()[i];
    const LN<N> z0 = f.load_half(AG_Z1, 0);
    if (E::HALVES == 1) {
        if (mpw_is_zero<N>(z0.v)) { for (int i = 0; i < 2 * FB; i++) out[i] = 0; return; }
        const LN<N> zi = sat_fp_mul<C>(mp_inv_plain<C>(z0), r2);               // R / Z
        agg_write_be<C>(out, sat_fp_mul<C>(f.load_half(AG_X1, 0), zi));
        agg_write_be<C>(out + FB, sat_fp_mul<C>(f.load_half(AG_Y1, 0), zi));
    } else {
        const LN<N> z1 = f.load_half(AG_Z1, 1);
        if (mpw_is_zero<N>(z0.v) && mpw_is_zero<N>(z1.v)) { for (int i = 0; i < 4 * FB; i++) out[i] = 0; return; }
        const LN<N> nrm = mp_add_f<C>(sat_fp_mul<C>(z0, z0), sat_fp_mul<C>(z1, z1));   // |Z|^2 / R
        const LN<N> ni = sat_fp_mul<C>(mp_inv_plain<C>(nrm), r2);              // R^2 / |Z|^2
        F2<C> zi;                                                              // R / Z
        zi.c0 = sat_fp_mul<C>(z0, ni);
        mp_neg<C>(zi.c1.v, sat_fp_mul<C>(z1, ni).v);
        for (int k = 0; k < 2; k++) {
            F2<C> a;
            a.c0 = f.load_half(AG_X1 + k, 0);
            a.c1 = f.load_half(AG_X1 + k, 1);
            const F2<C> x = sat_fp2_mul<C>(a, zi);
            agg_write_be<C>(out + (2 * k) * FB, x.c1);
            agg_write_be<C>(out + (2 * k + 1) * FB, x.c0);
        }
    }
}

// words of all levels / tickets of the cross-block tree over nb block values (fan-in fan), 3 HALVES N words per value
inline size_t agg_tree_values(size_t nb, size_t fan) {
    size_t w = 0;
    for (size_t c = nb; c > 1; c = (c + fan - 1) / fan) w += c;
    return w;
}
inline size_t agg_tree_tickets(size_t nb, size_t fan) {
    size_t k = 0;
    for (size_t c = nb; c > 1; c = (c + fan - 1) / fan) k += (c + fan - 1) / fan;
    return k;
}

#if defined(__CUDACC__)

// one addition on the six lanes of every active group of the warp; every lane of the warp must call it.  The only
// instance of the multiplication body in a kernel.
template <class E, class F> __device__ __noinline__ void agg_add(const F f, int role, bool act) {
#pragma unroll 1
    for (int step = 0; step < 3; step++) {
        if (step == 1 && !E::B3_MUL) {
            if (act) agg_b3_small<E>(f, role);
        } else if (act && agg_mul_active<E>(step, role)) {
            typename E::T a, b;
            agg_mul_operands<E>(f, step, role, a, b);
            f.store(agg_mul_dst(step, role), E::mul(a, b));
        }
        __syncwarp();
        if (act) agg_linear<E>(f, step, role);
        __syncwarp();
    }
}
// binary tree over the accumulators of the block's first `ngroups` groups; group 0 ends up with the sum
template <class E, int NGB, class F> __device__ __forceinline__ void agg_block_tree(const F& f, SvU4* slots, int role, bool lane_ok, int ngroups) {
    constexpr int W4 = F::W4;
#pragma unroll 1
    for (int st = 1; st < NGB; st <<= 1) {
        __syncthreads();
        if (st >= ngroups) break;
        const bool act = lane_ok && (f.q & (2 * st - 1)) == 0 && f.q + st < ngroups;
        if (act && role < 3 * E::HALVES) {
            const F o{slots, f.q + st};
            f.store_half(AG_X2 + role / E::HALVES, role % E::HALVES, o.load_half(AG_X1 + role / E::HALVES, role % E::HALVES));
        }
        __syncwarp();
        agg_add<E>(f, role, act);
        (void)W4;
    }
    __syncthreads();
}

// AggregatePoints: sum of n wire records -> one wire record.  `levels`: scratch of the cross-block tree
// (agg_tree_values(gridDim.x, NGB) values), `tickets`: zero-initialised, left zero.
// `trace` (optional, 8 words): globaltimer ns of the block that finishes -- its start, end of phases 1, 2, 3, 4.
template <class E, int WPB, int MINB = 1>
__global__ void __launch_bounds__(WPB * 32, MINB) k_agg(const uint8_t* __restrict__ pts, size_t n, uint32_t* __restrict__ levels,
                                                        unsigned* __restrict__ tickets, uint8_t* __restrict__ out,
                                                        unsigned long long* __restrict__ trace) {
    using C = typename E::Curve;
    constexpr int N = C::N, FB = C::FP_BYTES, NGB = WPB * AGG_GPW, NC = 2 * E::HALVES, VW = 3 * E::HALVES * N;
    constexpr size_t REC = (size_t)NC * FB;
    extern __shared__ uint4 agg_sm[];
    SvU4* slots = (SvU4*)agg_sm;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool lane_ok = lane < AGG_LANES * AGG_GPW;
    const int g = lane % AGG_GPW, role = lane / AGG_GPW;
    using F = AggFile<E, NGB>;
    const F f{slots, warp * AGG_GPW + g};
    unsigned long long tr[4] = {0, 0, 0, 0};
    if (trace) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tr[0]));
    if (lane_ok) agg_init_acc<E>(f, role);
    __syncwarp();
    // ---- phase 1: point i = it * (all groups) + (global group index): a warp reads five consecutive records
    const size_t total = (size_t)gridDim.x * NGB, mine = (size_t)blockIdx.x * NGB + f.q;
    const size_t warp_first = (size_t)blockIdx.x * NGB + (size_t)warp * AGG_GPW;
    constexpr unsigned GM = NC == 4 ? 0x8421u : 0x21u;     // lanes g + 5 c, c < NC
#pragma unroll 1
    for (size_t base = 0; base + warp_first < n; base += total) {
        const size_t i = base + mine;
        LN<N> v = agg_small<N>(0u);
        bool any = false, flag = false;
        if (lane_ok && role < NC && i < n) {
            const uint8_t* src = pts + i * REC + (size_t)role * FB;
            v = agg_read_be<C>(src);
            any = !mpw_is_zero<N>(v.v);
            flag = !C::IS_BN && role == 0 && (src[0] & 0x40);
        }
        const unsigned m_any = __ballot_sync(0xFFFFFFFFu, any), m_flag = __ballot_sync(0xFFFFFFFFu, flag);
        const bool inf = (m_any & (GM << g)) == 0 || ((m_flag >> g) & 1u);
        if (lane_ok) agg_store_input<E>(f, role, v, inf);
        __syncwarp();
        agg_add<E>(f, role, lane_ok);
    }
    if (trace) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tr[1]));
    // ---- phase 2: the block's groups
    const size_t first = (size_t)blockIdx.x * NGB;
    const int ngroups = n > first ? (int)(n - first < (size_t)NGB ? n - first : (size_t)NGB) : 1;   // groups that hold points
    agg_block_tree<E, NGB>(f, slots, role, lane_ok, ngroups);
    if (trace) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tr[2]));
    // ---- phase 3: cross-block tree (the last block of every NGB to arrive goes on)
    __shared__ int s_last;
    size_t count = gridDim.x, idx = blockIdx.x;
    uint32_t* lvl = levels;
    unsigned* cnt = tickets;
    const F f0{slots, 0};
#pragma unroll 1
    while (count > 1) {
        if (threadIdx.x < 32 && g == 0 && lane_ok) agg_export<E>(f0, role, AG_X1, lvl + idx * VW);
        __threadfence();
        __syncthreads();
        const size_t grp = idx / NGB;
        const int size = (int)(count - grp * NGB < (size_t)NGB ? count - grp * NGB : (size_t)NGB);
        if (threadIdx.x == 0) {
            const unsigned t = atomicAdd(cnt + grp, 1u);
            s_last = t == (unsigned)(size - 1);
            if (s_last) cnt[grp] = 0;
        }
        __syncthreads();
        if (!s_last) return;
        __threadfence();
        if (lane_ok && f.q < size) agg_import<E>(f, role, AG_X1, lvl + (grp * NGB + f.q) * VW);
        agg_block_tree<E, NGB>(f, slots, role, lane_ok, size);
        const size_t groups = (count + NGB - 1) / NGB;
        lvl += count * VW;
        cnt += groups;
        idx = grp;
        count = groups;
    }
    // ---- phase 4
    if (trace) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tr[3]));
    if (threadIdx.x == 0) {
        agg_store_affine<E>(f0, out);
        if (trace) {
            trace[0] = tr[0]; trace[1] = tr[1]; trace[2] = tr[2]; trace[3] = tr[3];
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(trace[4]));
        }
    }
}

#endif

}  // namespace bgls
