// Saturated-limb arithmetic with lazy reduction for the slot engine (slotvm.cuh).
//
// Fp elements are N 32-bit limbs in Montgomery form (R = 2^(32 N)), always canonical in [0, p).  The expensive part
// of every Fp2 operation is split into full products (N^2 IMAD.WIDE.U32 each) and Montgomery reductions
// (N^2 + N each): an Fp2 multiplication is 3 products + 2 reductions (Karatsuba on the unreduced double-width
// values), a squaring 2 + 2 -- against 3 x (2 N^2 + N) for three complete Montgomery multiplications.
//
// Products accumulate in two register files X (even-aligned 64-bit columns) and Y (odd-aligned), so that every
// 32x32+64 step is ONE IMAD.WIDE.U32 with its carry out in a predicate, consumed at once by an IADD3.X into a per-column
// carry counter: no carry chain links two products, which keeps a lone warp on a sub-partition issuing (measured
// on B200, tools/sat_bench.cu: 14.7 G Fp2 mul/s with one warp per sub-partition against 12.0 for row-chained carries
// and 10.4 for three CIOS multiplications, profiles/r2_b_*).
//
// This is the arithmetic the reference delegates to bn256.Pair / bls12.GT.Pair
// (/root/reference/curves/altbn128.go:136, curves/bls12_381.go:231).
#pragma once
#include "arith.cuh"
#include "curve_params.cuh"

namespace bgls {

// SAT_NI: bodies inlined into the one Fp2 multiplication of the interpreter (the hot 16 KB); SAT_COLD: real
// functions for the rarely used conversions, so that the whole kernel stays below the 32 KB of the L1.5 instruction
// cache (fully inlined everywhere it was 86 KB / 148 KB and every operation of a lone warp waited on instruction fetch)
#if defined(__CUDACC__)
#define SAT_NI __device__ __forceinline__
#define SAT_COLD __device__ __noinline__
#else
#define SAT_NI inline
#define SAT_COLD inline
#endif

template <int N> struct LN { uint32_t v[N]; };   // N limbs passed / returned by value (registers)

// (hi:lo) += a*b ; cnt += carry out   -- one IMAD.WIDE.U32 with predicate carry + one IADD3.X, no chain
HD void mad_wide_cnt(uint32_t& lo, uint32_t& hi, uint32_t& cnt, uint32_t a, uint32_t b) {
#if defined(__CUDACC__)
    asm volatile("mad.lo.cc.u32 %0, %3, %4, %0; madc.hi.cc.u32 %1, %3, %4, %1; addc.u32 %2, %2, 0;"
                 : "+r"(lo), "+r"(hi), "+r"(cnt) : "r"(a), "r"(b));
#else
    mad_wide_cc(lo, hi, a, b);
    addc(cnt, cnt, 0);
#endif
}

// T[0..2N) = a * b
template <int N> HD void mp_mul(uint32_t* T, const uint32_t* a, const uint32_t* b) {
    static_assert(N % 4 == 0, "limb count must be a multiple of four");
    uint32_t X[2 * N], Y[2 * N], Cn[2 * N + 2];   // value = X + (Y << 32) + Cn (Cn[k]: carries of weight k)
#pragma unroll
    for (int k = 0; k < 2 * N; k++) { X[k] = 0; Y[k] = 0; }
#pragma unroll
    for (int k = 0; k < 2 * N + 2; k++) Cn[k] = 0;
#pragma unroll
    for (int i = 0; i < N; i++) {
        const uint32_t bi = b[i];
#pragma unroll
        for (int j = 0; j < N; j++) {
            const int w = i + j;   // weight: even -> X pair (w, w+1), odd -> Y pair (w-1, w)
            if (i == 0) {
                if ((w & 1) == 0) mul_wide(X[w], X[w + 1], a[j], bi);
                else mul_wide(Y[w - 1], Y[w], a[j], bi);
            } else if ((w & 1) == 0) mad_wide_cnt(X[w], X[w + 1], Cn[w + 2], a[j], bi);
            else mad_wide_cnt(Y[w - 1], Y[w], Cn[w + 2], a[j], bi);
        }
    }
    T[0] = X[0];
    add_cc(T[1], X[1], Y[0]);
#pragma unroll
    for (int k = 2; k < 2 * N - 1; k++) addc_cc(T[k], X[k], Y[k - 1]);
    addc(T[2 * N - 1], X[2 * N - 1], Y[2 * N - 2]);
    add_cc(T[2], T[2], Cn[2]);
#pragma unroll
    for (int k = 3; k < 2 * N - 1; k++) addc_cc(T[k], T[k], Cn[k]);
    addc(T[2 * N - 1], T[2 * N - 1], Cn[2 * N - 1]);
}

// Montgomery reduction: r = T / R mod p for T < p R (2N words), r canonical.  A is destroyed.
// Row i adds m_i p 2^(32 i) with the same even/odd split and carry counters (K[k]: weight k).
template <class C> HD void mp_redc(uint32_t* r, uint32_t* A) {
    constexpr int N = C::N;
    uint32_t B[2 * N], K[2 * N + 2];   // B odd-aligned: B[k] has weight k + 1
#pragma unroll
    for (int k = 0; k < 2 * N; k++) B[k] = 0;
#pragma unroll
    for (int k = 0; k < 2 * N + 2; k++) K[k] = 0;
    uint32_t c = 0;                    // carry into weight i from the words already cancelled
#pragma unroll
    for (int i = 0; i < N; i++) {
        const uint32_t lo = A[i] + (i ? B[i - 1] : 0u) + K[i] + c;
        const uint32_t m = lo * C::N0;
#pragma unroll
        for (int j = 0; j < N; j++) {
            const int w = i + j;
            if ((w & 1) == 0) mad_wide_cnt(A[w], A[w + 1], K[w + 2], C::p(j), m);
            else mad_wide_cnt(B[w - 1], B[w], K[w + 2], C::p(j), m);
        }
        // the word of weight i is now 0 mod 2^32: its carry moves on
        const unsigned long long s = (unsigned long long)A[i] + (i ? B[i - 1] : 0u) + K[i] + c;
        c = (uint32_t)(s >> 32);
    }
    // t = (A + (B << 32) + K) >> 32N  + c        (t < 2p < 2^(32N))
    uint32_t t[N];
    add_cc(t[0], A[N], B[N - 1]);
#pragma unroll
    for (int k = 1; k < N; k++) addc_cc(t[k], A[N + k], B[N + k - 1]);
    add_cc(t[0], t[0], c);
#pragma unroll
    for (int k = 1; k < N; k++) addc_cc(t[k], t[k], 0);
    add_cc(t[0], t[0], K[N]);
#pragma unroll
    for (int k = 1; k < N; k++) addc_cc(t[k], t[k], K[N + k]);
    // conditional subtraction of p
    uint32_t d[N], br;
    sub_cc(d[0], t[0], C::p(0));
#pragma unroll
    for (int k = 1; k < N; k++) subc_cc(d[k], t[k], C::p(k));
    subc(br, 0, 0);
#pragma unroll
    for (int k = 0; k < N; k++) r[k] = br ? t[k] : d[k];
}

// ---- canonical linear operations on raw limb arrays
template <class C> HD void mp_add(uint32_t* r, const uint32_t* a, const uint32_t* b) {   // (a + b) mod p
    constexpr int N = C::N;
    uint32_t t[N], d[N], br;
    add_cc(t[0], a[0], b[0]);
#pragma unroll
    for (int i = 1; i < N; i++) addc_cc(t[i], a[i], b[i]);
    // 2p < 2^(32N): no carry out
    sub_cc(d[0], t[0], C::p(0));
#pragma unroll
    for (int i = 1; i < N; i++) subc_cc(d[i], t[i], C::p(i));
    subc(br, 0, 0);
#pragma unroll
    for (int i = 0; i < N; i++) r[i] = br ? t[i] : d[i];
}
template <class C> HD void mp_sub(uint32_t* r, const uint32_t* a, const uint32_t* b) {   // (a - b) mod p
    constexpr int N = C::N;
    uint32_t t[N], br;
    sub_cc(t[0], a[0], b[0]);
#pragma unroll
    for (int i = 1; i < N; i++) subc_cc(t[i], a[i], b[i]);
    subc(br, 0, 0);
    add_cc(r[0], t[0], C::p(0) & br);
#pragma unroll
    for (int i = 1; i < N - 1; i++) addc_cc(r[i], t[i], C::p(i) & br);
    addc(r[N - 1], t[N - 1], C::p(N - 1) & br);
}
// a + b without reduction (caller guarantees a + b < 2^(32N))
template <int N> HD void mp_add_nr(uint32_t* r, const uint32_t* a, const uint32_t* b) {
    add_cc(r[0], a[0], b[0]);
#pragma unroll
    for (int i = 1; i < N - 1; i++) addc_cc(r[i], a[i], b[i]);
    addc(r[N - 1], a[N - 1], b[N - 1]);
}
template <class C> HD void mp_half(uint32_t* r, const uint32_t* a) {   // a / 2 mod p
    constexpr int N = C::N;
    const uint32_t odd = 0u - (a[0] & 1u);
    uint32_t t[N];
    add_cc(t[0], a[0], C::p(0) & odd);
#pragma unroll
    for (int i = 1; i < N - 1; i++) addc_cc(t[i], a[i], C::p(i) & odd);
    addc(t[N - 1], a[N - 1], C::p(N - 1) & odd);   // a + p < 2^(32N)
#pragma unroll
    for (int i = 0; i < N - 1; i++) r[i] = (t[i] >> 1) | (t[i + 1] << 31);
    r[N - 1] = t[N - 1] >> 1;
}
template <class C> HD void mp_neg(uint32_t* r, const uint32_t* a) {   // p - a, 0 stays 0
    constexpr int N = C::N;
    uint32_t any = 0;
#pragma unroll
    for (int i = 0; i < N; i++) any |= a[i];
    const uint32_t m = any ? 0xFFFFFFFFu : 0u;
    sub_cc(r[0], C::p(0) & m, a[0]);
#pragma unroll
    for (int i = 1; i < N - 1; i++) subc_cc(r[i], C::p(i) & m, a[i]);
    subc(r[N - 1], C::p(N - 1) & m, a[N - 1]);
}
// 2N-word subtraction; *borrow_mask = 0xffffffff when a < b
template <int N> HD void wide_sub(uint32_t* r, const uint32_t* a, const uint32_t* b, uint32_t* borrow_mask) {
    sub_cc(r[0], a[0], b[0]);
#pragma unroll
    for (int i = 1; i < 2 * N; i++) subc_cc(r[i], a[i], b[i]);
    subc(*borrow_mask, 0, 0);
}

// ---- the non-inlined bodies (operands and results in registers)
template <int N> SAT_NI LN<2 * N> mp_mul_f(LN<N> a, LN<N> b) {
    LN<2 * N> t;
    mp_mul<N>(t.v, a.v, b.v);
    return t;
}
template <class C> SAT_NI LN<C::N> mp_redc_f(LN<2 * C::N> t) {
    LN<C::N> r;
    mp_redc<C>(r.v, t.v);
    return r;
}
template <class C> SAT_NI LN<C::N> mp_add_f(LN<C::N> a, LN<C::N> b) {
    LN<C::N> r;
    mp_add<C>(r.v, a.v, b.v);
    return r;
}
template <class C> SAT_NI LN<C::N> mp_sub_f(LN<C::N> a, LN<C::N> b) {
    LN<C::N> r;
    mp_sub<C>(r.v, a.v, b.v);
    return r;
}

// ---- Fp2 on (c0, c1) limb structs; results canonical
template <class C> struct F2 { LN<C::N> c0, c1; };

// r = a * b : Karatsuba on the unreduced products, two reductions
template <class C> HD F2<C> sat_fp2_mul(const F2<C>& a, const F2<C>& b) {
    constexpr int N = C::N;
    LN<N> sa, sb;
    mp_add_nr<N>(sa.v, a.c0.v, a.c1.v);          // < 2p, fits
    mp_add_nr<N>(sb.v, b.c0.v, b.c1.v);
    LN<2 * N> T0 = mp_mul_f<N>(a.c0, b.c0);
    LN<2 * N> T1 = mp_mul_f<N>(a.c1, b.c1);
    LN<2 * N> T2 = mp_mul_f<N>(sa, sb);          // < 4 p^2 < 2^(64N)
    uint32_t bm;
    // im = T2 - T0 - T1 = a0 b1 + a1 b0 in [0, 2 p^2)
    wide_sub<N>(T2.v, T2.v, T0.v, &bm);
    wide_sub<N>(T2.v, T2.v, T1.v, &bm);
    // re = T0 - T1 in (-p^2, p^2): add p R when negative
    wide_sub<N>(T0.v, T0.v, T1.v, &bm);
    add_cc(T0.v[N], T0.v[N], C::p(0) & bm);
#pragma unroll
    for (int i = 1; i < N - 1; i++) addc_cc(T0.v[N + i], T0.v[N + i], C::p(i) & bm);
    addc(T0.v[2 * N - 1], T0.v[2 * N - 1], C::p(N - 1) & bm);
    F2<C> r;
    r.c0 = mp_redc_f<C>(T0);
    r.c1 = mp_redc_f<C>(T2);
    return r;
}
// r = a^2 : re = (a0 + a1)(a0 - a1), im = (2 a0) a1
template <class C> HD F2<C> sat_fp2_sqr(const F2<C>& a) {
    constexpr int N = C::N;
    LN<N> s, e;
    mp_add_nr<N>(s.v, a.c0.v, a.c1.v);           // < 2p
    LN<N> d = mp_sub_f<C>(a.c0, a.c1);           // canonical
    mp_add_nr<N>(e.v, a.c0.v, a.c0.v);           // < 2p
    F2<C> r;
    r.c0 = mp_redc_f<C>(mp_mul_f<N>(s, d));      // < 2 p^2
    r.c1 = mp_redc_f<C>(mp_mul_f<N>(e, a.c1));   // < 2 p^2
    return r;
}
// ---- xi = 9 + i on altbn128: r = 9 x + y or 9 x - y mod p in ONE reduction
// multiples k p (k = 0..11) of the altbn128 prime, nine words padded to twelve (three 16-byte loads)
#if defined(__CUDACC__)
static __device__ const uint32_t BN254_KP[12][12] __attribute__((aligned(16))) = {
#else
static const uint32_t BN254_KP[12][12] __attribute__((aligned(16))) = {
#endif
    {0x00000000u, 0x00000000u, 0x00000000u, 0x00000000u, 0x00000000u, 0x00000000u, 0x00000000u, 0x00000000u, 0x00000000u, 0x00000000u, 0x00000000u, 0x00000000u},
    {0xd87cfd47u, 0x3c208c16u, 0x6871ca8du, 0x97816a91u, 0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u, 0x00000000u, 0x00000000u, 0x00000000u, 0x00000000u},
    {0xb0f9fa8eu, 0x7841182du, 0xd0e3951au, 0x2f02d522u, 0x0302b0bbu, 0x70a08b6du, 0xc2634053u, 0x60c89ce5u, 0x00000000u, 0x00000000u, 0x00000000u, 0x00000000u},
    {0x8976f7d5u, 0xb461a444u, 0x39555fa7u, 0xc6843fb4u, 0x84840918u, 0x28f0d123u, 0xa394e07du, 0x912ceb58u, 0x00000000u, 0x00000000u, 0x00000000u, 0x00000000u},
    {0x61f3f51cu, 0xf082305bu, 0xa1c72a34u, 0x5e05aa45u, 0x06056176u, 0xe14116dau, 0x84c680a6u, 0xc19139cbu, 0x00000000u, 0x00000000u, 0x00000000u, 0x00000000u},
    {0x3a70f263u, 0x2ca2bc72u, 0x0a38f4c2u, 0xf58714d7u, 0x8786b9d3u, 0x99915c90u, 0x65f820d0u, 0xf1f5883eu, 0x00000000u, 0x00000000u, 0x00000000u, 0x00000000u},
    {0x12edefaau, 0x68c34889u, 0x72aabf4fu, 0x8d087f68u, 0x09081231u, 0x51e1a247u, 0x4729c0fau, 0x2259d6b1u, 0x00000001u, 0x00000000u, 0x00000000u, 0x00000000u},
    {0xeb6aecf1u, 0xa4e3d49fu, 0xdb1c89dcu, 0x2489e9f9u, 0x8a896a8fu, 0x0a31e7fdu, 0x285b6124u, 0x52be2524u, 0x00000001u, 0x00000000u, 0x00000000u, 0x00000000u},
    {0xc3e7ea38u, 0xe10460b6u, 0x438e5469u, 0xbc0b548bu, 0x0c0ac2ecu, 0xc2822db4u, 0x098d014du, 0x83227397u, 0x00000001u, 0x00000000u, 0x00000000u, 0x00000000u},
    {0x9c64e77fu, 0x1d24eccdu, 0xac001ef7u, 0x538cbf1cu, 0x8d8c1b4au, 0x7ad2736au, 0xeabea177u, 0xb386c209u, 0x00000001u, 0x00000000u, 0x00000000u, 0x00000000u},
    {0x74e1e4c6u, 0x594578e4u, 0x1471e984u, 0xeb0e29aeu, 0x0f0d73a7u, 0x3322b921u, 0xcbf041a1u, 0xe3eb107cu, 0x00000001u, 0x00000000u, 0x00000000u, 0x00000000u},
    {0x4d5ee20du, 0x956604fbu, 0x7ce3b411u, 0x828f943fu, 0x908ecc05u, 0xeb72fed7u, 0xad21e1cau, 0x144f5eefu, 0x00000002u, 0x00000000u, 0x00000000u, 0x00000000u}};
// t = (x << 3) + x + (y | p - y) in nine words (< 11 p); the quotient floor(t / p) is estimated from the two top words
// (never too large, at most one too small: T / (P7 + 1) <= t / p with T = t >> 224, P7 = p >> 224, and the truncations
// lose less than 2^-25), the table row q p is subtracted and one conditional subtraction finishes.  ~75 instructions
// against ~125 for three canonical doublings, an addition and an addition / subtraction.
template <class C> HD LN<8> bn_lin9(const LN<8>& x, const LN<8>& y, bool sub) {
    constexpr int N = 8;
    static_assert(C::N == N && C::IS_BN, "altbn128 only");
    uint32_t w[N], s[N + 1], t[N + 1];
    if (sub) {                                   // p - y in [1, p]
        sub_cc(w[0], C::p(0), y.v[0]);
#pragma unroll
        for (int i = 1; i < N - 1; i++) subc_cc(w[i], C::p(i), y.v[i]);
        subc(w[N - 1], C::p(N - 1), y.v[N - 1]);
    } else {
#pragma unroll
        for (int i = 0; i < N; i++) w[i] = y.v[i];
    }
    s[0] = x.v[0] << 3;
#pragma unroll
    for (int i = 1; i < N; i++) s[i] = (x.v[i] << 3) | (x.v[i - 1] >> 29);
    s[N] = x.v[N - 1] >> 29;
    add_cc(t[0], s[0], x.v[0]);
#pragma unroll
    for (int i = 1; i < N; i++) addc_cc(t[i], s[i], x.v[i]);
    addc(t[N], s[N], 0);
    add_cc(t[0], t[0], w[0]);
#pragma unroll
    for (int i = 1; i < N; i++) addc_cc(t[i], t[i], w[i]);
    addc(t[N], t[N], 0);
    const unsigned long long T = ((unsigned long long)t[N] << 32) | t[N - 1];
    const uint32_t q = (uint32_t)(((T >> 3) * 2840127684ull) >> 58);   // floor(2^61 / (P7 + 1)) = 2840127684
    uint32_t kp[12];
#if defined(__CUDA_ARCH__)
    {
        const uint4* row = (const uint4*)BN254_KP[q];
        const uint4 a = row[0], b = row[1], c = row[2];
        kp[0] = a.x; kp[1] = a.y; kp[2] = a.z; kp[3] = a.w; kp[4] = b.x; kp[5] = b.y; kp[6] = b.z; kp[7] = b.w; kp[8] = c.x;
    }
#else
    for (int i = 0; i < 9; i++) kp[i] = BN254_KP[q][i];
#endif
    sub_cc(t[0], t[0], kp[0]);
#pragma unroll
    for (int i = 1; i < N; i++) subc_cc(t[i], t[i], kp[i]);
    subc(t[N], t[N], kp[N]);                     // t in [0, 2p): the top word is 0 now
    uint32_t d[N], br;
    sub_cc(d[0], t[0], C::p(0));
#pragma unroll
    for (int i = 1; i < N; i++) subc_cc(d[i], t[i], C::p(i));
    subc(br, 0, 0);
    LN<N> r;
#pragma unroll
    for (int i = 0; i < N; i++) r.v[i] = br ? t[i] : d[i];
    return r;
}
template <class C, bool BN = C::IS_BN> struct SatXi;
template <class C> struct SatXi<C, true> {       // (9 + i)(a0 + a1 i) = (9 a0 - a1) + (9 a1 + a0) i
    static HD void mul(LN<C::N>& r0, LN<C::N>& r1, const LN<C::N>& a0, const LN<C::N>& a1) {
        r0 = bn_lin9<C>(a0, a1, true);
        r1 = bn_lin9<C>(a1, a0, false);
    }
};
template <class C> struct SatXi<C, false> {      // (1 + i)(a0 + a1 i) = (a0 - a1) + (a0 + a1) i
    static HD void mul(LN<C::N>& r0, LN<C::N>& r1, const LN<C::N>& a0, const LN<C::N>& a1) {
        r0 = mp_sub_f<C>(a0, a1);
        r1 = mp_add_f<C>(a0, a1);
    }
};

// Fp Montgomery product (conversion in / out of Montgomery form)
template <class C> SAT_COLD LN<C::N> sat_fp_mul(LN<C::N> a, LN<C::N> b) { return mp_redc_f<C>(mp_mul_f<C::N>(a, b)); }
// r = xi * a
template <class C> HD F2<C> sat_fp2_mul_xi(const F2<C>& a) {
    F2<C> r;
    SatXi<C>::mul(r.c0, r.c1, a.c0, a.c1);
    return r;
}

}  // namespace bgls
