// Fp / Fp2 / Fp6 / Fp12 towers over 32-bit Montgomery limbs, generic in the curve traits C
// (BN254: N = 8, BLS381: N = 12; see curve_params.cuh).
//
//   Fp2 = Fp[i]/(i^2+1),  Fp6 = Fp2[v]/(v^3 - xi),  Fp12 = Fp6[w]/(w^2 - v)
//   xi = 9+i (altbn128), 1+i (bls12-381)
//
// This is the arithmetic the reference delegates to bn256.Pair / bls12.GT.Pair
// (/root/reference/curves/altbn128.go:136, curves/bls12_381.go:231) and to G1/G2/GT Add
// (altbn128.go:59-66,181-188,264-271; bls12_381.go:33-41,94-102,160-168).
// All values are kept fully reduced in [0, p).
#pragma once
#include "arith.cuh"
#include "curve_params.cuh"
#include "inv.cuh"

namespace bgls {

template <class C> struct Fp { uint32_t v[C::N]; };
template <class C> struct Fp2 { Fp<C> c0, c1; };
template <class C> struct Fp6 { Fp2<C> a0, a1, a2; };
template <class C> struct Fp12 { Fp6<C> c0, c1; };

// ------------------------------------------------------------------------------------ Fp
template <class C> HD void fp_set(Fp<C>& r, const uint32_t* s) {
#pragma unroll
    for (int i = 0; i < C::N; i++) r.v[i] = s[i];
}
template <class C> HD void fp_zero(Fp<C>& r) {
#pragma unroll
    for (int i = 0; i < C::N; i++) r.v[i] = 0;
}
template <class C> HD bool fp_is_zero(const Fp<C>& a) {
    uint32_t x = 0;
#pragma unroll
    for (int i = 0; i < C::N; i++) x |= a.v[i];
    return x == 0;
}
template <class C> HD bool fp_eq(const Fp<C>& a, const Fp<C>& b) {
    uint32_t x = 0;
#pragma unroll
    for (int i = 0; i < C::N; i++) x |= a.v[i] ^ b.v[i];
    return x == 0;
}
// r = t - p if t >= p (t < 2p, top carry word `hi` may be 1)
template <class C> HD void fp_final_sub(Fp<C>& r, const uint32_t* t, uint32_t hi) {
    constexpr int N = C::N;
    uint32_t s[N], b;
    sub_cc(s[0], t[0], C::p(0));
#pragma unroll
    for (int i = 1; i < N; i++) subc_cc(s[i], t[i], C::p(i));
    subc(b, hi, 0);  // b == 0 -> no borrow -> take s
#pragma unroll
    for (int i = 0; i < N; i++) r.v[i] = b ? t[i] : s[i];
}
template <class C> HD void fp_add(Fp<C>& r, const Fp<C>& a, const Fp<C>& b) {
    constexpr int N = C::N;
    uint32_t t[N], hi;
    add_cc(t[0], a.v[0], b.v[0]);
#pragma unroll
    for (int i = 1; i < N; i++) addc_cc(t[i], a.v[i], b.v[i]);
    addc(hi, 0, 0);
    fp_final_sub<C>(r, t, hi);
}
template <class C> HD void fp_sub(Fp<C>& r, const Fp<C>& a, const Fp<C>& b) {
    constexpr int N = C::N;
    uint32_t t[N], br;
    sub_cc(t[0], a.v[0], b.v[0]);
#pragma unroll
    for (int i = 1; i < N; i++) subc_cc(t[i], a.v[i], b.v[i]);
    subc(br, 0, 0);  // 0xffffffff if borrow
    add_cc(r.v[0], t[0], C::p(0) & br);
#pragma unroll
    for (int i = 1; i < N - 1; i++) addc_cc(r.v[i], t[i], C::p(i) & br);
    addc(r.v[N - 1], t[N - 1], C::p(N - 1) & br);
}
template <class C> HD void fp_neg(Fp<C>& r, const Fp<C>& a) {
    Fp<C> z;
    fp_zero(z);
    fp_sub(r, z, a);
}
template <class C> HD void fp_dbl(Fp<C>& r, const Fp<C>& a) { fp_add(r, a, a); }

// Montgomery product a*b/R mod p.  Even/odd column accumulators so that every 32x32+64 step is
// one IMAD.WIDE.U32 with the carry riding in a predicate (2N^2 + N multiplies for N limbs).
template <class C> HD void fp_mul(Fp<C>& r, const Fp<C>& a, const Fp<C>& b) {
    constexpr int N = C::N;
    uint32_t X[N], Y[N];
    // ---- i = 0
    {
        const uint32_t bi = b.v[0];
#pragma unroll
        for (int j = 0; j < N; j += 2) mul_wide(X[j], X[j + 1], a.v[j], bi);
#pragma unroll
        for (int j = 1; j < N; j += 2) mul_wide(Y[j - 1], Y[j], a.v[j], bi);
        const uint32_t m = X[0] * C::N0;
        mad_wide_cc(Y[0], Y[1], C::p(1), m);
#pragma unroll
        for (int j = 3; j < N; j += 2) madc_wide_cc(Y[j - 1], Y[j], C::p(j), m);
        mad_wide_cc(X[0], X[1], C::p(0), m);
#pragma unroll
        for (int j = 2; j < N; j += 2) madc_wide_cc(X[j], X[j + 1], C::p(j), m);
        addc(Y[N - 1], Y[N - 1], 0);
    }
#pragma unroll
    for (int i = 1; i < N; i++) {
        // E: even-aligned accumulator for this step, Z: the previous one (Z[0] == 0), shifted by two words
        uint32_t* E = (i & 1) ? Y : X;
        uint32_t* Z = (i & 1) ? X : Y;
        const uint32_t bi = b.v[i];
        add_cc(E[0], E[0], Z[1]);
#pragma unroll
        for (int j = 1; j < N - 1; j += 2) madc_wide_cc3(Z[j - 1], Z[j], a.v[j], bi, Z[j + 1], Z[j + 2]);
        madc_wide_cc3(Z[N - 2], Z[N - 1], a.v[N - 1], bi, 0, 0);
        mad_wide_cc(E[0], E[1], a.v[0], bi);
#pragma unroll
        for (int j = 2; j < N; j += 2) madc_wide_cc(E[j], E[j + 1], a.v[j], bi);
        addc(Z[N - 1], Z[N - 1], 0);
        const uint32_t m = E[0] * C::N0;
        mad_wide_cc(Z[0], Z[1], C::p(1), m);
#pragma unroll
        for (int j = 3; j < N; j += 2) madc_wide_cc(Z[j - 1], Z[j], C::p(j), m);
        mad_wide_cc(E[0], E[1], C::p(0), m);
#pragma unroll
        for (int j = 2; j < N; j += 2) madc_wide_cc(E[j], E[j + 1], C::p(j), m);
        addc(Z[N - 1], Z[N - 1], 0);
    }
    // N even: the last step had E = Y (now Y[0] == 0, to be shifted one word) and Z = X
    uint32_t t[N];
    add_cc(t[0], X[0], Y[1]);
#pragma unroll
    for (int k = 1; k < N - 1; k++) addc_cc(t[k], X[k], Y[k + 1]);
    addc(t[N - 1], X[N - 1], 0);
    fp_final_sub<C>(r, t, 0);
}
template <class C> HD void fp_sqr(Fp<C>& r, const Fp<C>& a) { fp_mul(r, a, a); }

// a^e, e = nl 32-bit little-endian limbs (public exponent: every thread takes the same branches).
// Fixed 4-bit windows: 14 multiplications for the table, then 4 squarings and at most one multiplication per nibble
// (the exponents here are (q-2), (q+1)/4, (q-1)/2: about a quarter fewer multiplications than bit by bit).
template <class C> HDNI void fp_pow(Fp<C>& r, const Fp<C>& a, const uint32_t* e, int nl) {
    Fp<C> tab[16];
    fp_set(tab[0], C::R1());
    tab[1] = a;
    for (int k = 2; k < 16; k++) fp_mul(tab[k], tab[k - 1], a);
    Fp<C> acc;
    bool started = false;
    for (int i = nl * 8 - 1; i >= 0; i--) {
        const uint32_t nib = (e[i >> 3] >> (4 * (i & 7))) & 0xFu;
        if (started) {
            fp_sqr(acc, acc);
            fp_sqr(acc, acc);
            fp_sqr(acc, acc);
            fp_sqr(acc, acc);
            if (nib) fp_mul(acc, acc, tab[nib]);
        } else if (nib) {
            acc = tab[nib];
            started = true;
        }
    }
    if (!started) fp_set(acc, C::R1());
    r = acc;
}
// r = a^-1 (Montgomery form in and out; a = 0 gives 0 as the Fermat power did).  Binary inversion of the limb value
// a R (inv.cuh), the division by 2^k as two Montgomery products, then back to Montgomery form: (a R)^-1 R^2 = a^-1 R.
template <class C> HD void fp_inv(Fp<C>& r, const Fp<C>& a) {
    if (fp_is_zero(a)) { fp_zero(r); return; }
    Fp<C> x, e1, e2, r2;
    const int k = mp_almost_inv<C>(x.v, a.v);
    inv_shift_limbs<C::N>(e1.v, e2.v, k);
    fp_set(r2, C::R2());
    fp_mul(x, x, e1);
    fp_mul(x, x, e2);
    fp_mul(x, x, r2);
    fp_mul(r, x, r2);
}

// big-endian bytes (reference wire layout) <-> Montgomery limbs
template <class C> HD void fp_from_be(Fp<C>& r, const uint8_t* be) {
    constexpr int N = C::N;
    Fp<C> t, r2;
#pragma unroll
    for (int i = 0; i < N; i++) {
        const uint8_t* q = be + 4 * (N - 1 - i);
        t.v[i] = ((uint32_t)q[0] << 24) | ((uint32_t)q[1] << 16) | ((uint32_t)q[2] << 8) | q[3];
    }
    fp_set(r2, C::R2());
    fp_mul(r, t, r2);
}
template <class C> HD void fp_to_be(uint8_t* be, const Fp<C>& a) {
    constexpr int N = C::N;
    Fp<C> one, t;
    fp_zero(one);
    one.v[0] = 1;
    fp_mul(t, a, one);
#pragma unroll
    for (int i = 0; i < N; i++) {
        uint8_t* q = be + 4 * (N - 1 - i);
        q[0] = (uint8_t)(t.v[i] >> 24);
        q[1] = (uint8_t)(t.v[i] >> 16);
        q[2] = (uint8_t)(t.v[i] >> 8);
        q[3] = (uint8_t)t.v[i];
    }
}

// ------------------------------------------------------------------------------------ Fp2
template <class C> HD void fp2_add(Fp2<C>& r, const Fp2<C>& a, const Fp2<C>& b) { fp_add(r.c0, a.c0, b.c0); fp_add(r.c1, a.c1, b.c1); }
template <class C> HD void fp2_sub(Fp2<C>& r, const Fp2<C>& a, const Fp2<C>& b) { fp_sub(r.c0, a.c0, b.c0); fp_sub(r.c1, a.c1, b.c1); }
template <class C> HD void fp2_neg(Fp2<C>& r, const Fp2<C>& a) { fp_neg(r.c0, a.c0); fp_neg(r.c1, a.c1); }
template <class C> HD void fp2_dbl(Fp2<C>& r, const Fp2<C>& a) { fp_dbl(r.c0, a.c0); fp_dbl(r.c1, a.c1); }
template <class C> HD void fp2_conj(Fp2<C>& r, const Fp2<C>& a) { r.c0 = a.c0; fp_neg(r.c1, a.c1); }
template <class C> HD void fp2_zero(Fp2<C>& r) { fp_zero(r.c0); fp_zero(r.c1); }
template <class C> HD void fp2_one(Fp2<C>& r) { fp_set(r.c0, C::R1()); fp_zero(r.c1); }
template <class C> HD bool fp2_is_zero(const Fp2<C>& a) { return fp_is_zero(a.c0) && fp_is_zero(a.c1); }
template <class C> HD bool fp2_eq(const Fp2<C>& a, const Fp2<C>& b) { return fp_eq(a.c0, b.c0) && fp_eq(a.c1, b.c1); }
template <class C> HD void fp2_set(Fp2<C>& r, const uint32_t* s) { fp_set(r.c0, s); fp_set(r.c1, s + C::N); }

template <class C> HDNI void fp2_mul(Fp2<C>& r, const Fp2<C>& a, const Fp2<C>& b) {
    Fp<C> v0, v1, s, t;
    fp_mul(v0, a.c0, b.c0);
    fp_mul(v1, a.c1, b.c1);
    fp_add(s, a.c0, a.c1);
    fp_add(t, b.c0, b.c1);
    fp_mul(s, s, t);
    fp_sub(s, s, v0);
    fp_sub(r.c1, s, v1);
    fp_sub(r.c0, v0, v1);
}
template <class C> HDNI void fp2_sqr(Fp2<C>& r, const Fp2<C>& a) {
    Fp<C> s, d, m;
    fp_add(s, a.c0, a.c1);
    fp_sub(d, a.c0, a.c1);
    fp_mul(m, a.c0, a.c1);
    fp_mul(r.c0, s, d);
    fp_dbl(r.c1, m);
}
template <class C> HDNI void fp2_mul_fp(Fp2<C>& r, const Fp2<C>& a, const Fp<C>& b) {
    fp_mul(r.c0, a.c0, b);
    fp_mul(r.c1, a.c1, b);
}
// multiply by xi
template <class C> HD void fp2_mul_xi(Fp2<C>& r, const Fp2<C>& a) {
    Fp<C> n0, n1;
    if (C::IS_BN) {  // (9+i)(a0+a1 i) = (9a0 - a1) + (9a1 + a0) i
        Fp<C> t0, t1;
        fp_dbl(t0, a.c0); fp_dbl(t0, t0); fp_dbl(t0, t0); fp_add(t0, t0, a.c0);
        fp_dbl(t1, a.c1); fp_dbl(t1, t1); fp_dbl(t1, t1); fp_add(t1, t1, a.c1);
        fp_sub(n0, t0, a.c1);
        fp_add(n1, t1, a.c0);
    } else {  // (1+i)(a0+a1 i) = (a0 - a1) + (a0 + a1) i
        fp_sub(n0, a.c0, a.c1);
        fp_add(n1, a.c0, a.c1);
    }
    r.c0 = n0;
    r.c1 = n1;
}
template <class C> HDNI void fp2_inv(Fp2<C>& r, const Fp2<C>& a) {
    Fp<C> n, t;
    fp_sqr(n, a.c0);
    fp_sqr(t, a.c1);
    fp_add(n, n, t);
    fp_inv(n, n);
    fp_mul(r.c0, a.c0, n);
    fp_mul(t, a.c1, n);
    fp_neg(r.c1, t);
}

// ------------------------------------------------------------------------------------ Fp6
template <class C> HD void fp6_add(Fp6<C>& r, const Fp6<C>& a, const Fp6<C>& b) { fp2_add(r.a0, a.a0, b.a0); fp2_add(r.a1, a.a1, b.a1); fp2_add(r.a2, a.a2, b.a2); }
template <class C> HD void fp6_sub(Fp6<C>& r, const Fp6<C>& a, const Fp6<C>& b) { fp2_sub(r.a0, a.a0, b.a0); fp2_sub(r.a1, a.a1, b.a1); fp2_sub(r.a2, a.a2, b.a2); }
template <class C> HD void fp6_neg(Fp6<C>& r, const Fp6<C>& a) { fp2_neg(r.a0, a.a0); fp2_neg(r.a1, a.a1); fp2_neg(r.a2, a.a2); }
template <class C> HD void fp6_mul_v(Fp6<C>& r, const Fp6<C>& a) {
    Fp2<C> t;
    fp2_mul_xi(t, a.a2);
    r.a2 = a.a1;
    r.a1 = a.a0;
    r.a0 = t;
}
template <class C> HDNI void fp6_mul(Fp6<C>& r, const Fp6<C>& a, const Fp6<C>& b) {
    Fp2<C> v0, v1, v2, s, t, u;
    Fp6<C> o;
    fp2_mul(v0, a.a0, b.a0);
    fp2_mul(v1, a.a1, b.a1);
    fp2_mul(v2, a.a2, b.a2);
    fp2_add(s, a.a1, a.a2); fp2_add(t, b.a1, b.a2); fp2_mul(u, s, t);
    fp2_sub(u, u, v1); fp2_sub(u, u, v2); fp2_mul_xi(u, u); fp2_add(o.a0, v0, u);
    fp2_add(s, a.a0, a.a1); fp2_add(t, b.a0, b.a1); fp2_mul(u, s, t);
    fp2_sub(u, u, v0); fp2_sub(u, u, v1); fp2_mul_xi(s, v2); fp2_add(o.a1, u, s);
    fp2_add(s, a.a0, a.a2); fp2_add(t, b.a0, b.a2); fp2_mul(u, s, t);
    fp2_sub(u, u, v0); fp2_sub(u, u, v2); fp2_add(o.a2, u, v1);
    r = o;
}
// a * (b0 + b1 v)
template <class C> HDNI void fp6_mul_by_01(Fp6<C>& r, const Fp6<C>& a, const Fp2<C>& b0, const Fp2<C>& b1) {
    Fp2<C> v0, v1, s, t, u;
    Fp6<C> o;
    fp2_mul(v0, a.a0, b0);
    fp2_mul(v1, a.a1, b1);
    fp2_mul(u, a.a2, b1); fp2_mul_xi(u, u); fp2_add(o.a0, v0, u);
    fp2_add(s, a.a0, a.a1); fp2_add(t, b0, b1); fp2_mul(u, s, t);
    fp2_sub(u, u, v0); fp2_sub(o.a1, u, v1);
    fp2_mul(u, a.a2, b0); fp2_add(o.a2, u, v1);
    r = o;
}
template <class C> HDNI void fp6_mul_by_0(Fp6<C>& r, const Fp6<C>& a, const Fp2<C>& b0) {
    fp2_mul(r.a0, a.a0, b0);
    fp2_mul(r.a1, a.a1, b0);
    fp2_mul(r.a2, a.a2, b0);
}
// a * (b1 v)
template <class C> HDNI void fp6_mul_by_1(Fp6<C>& r, const Fp6<C>& a, const Fp2<C>& b1) {
    Fp2<C> t0, t1, t2;
    fp2_mul(t0, a.a2, b1); fp2_mul_xi(t0, t0);
    fp2_mul(t1, a.a0, b1);
    fp2_mul(t2, a.a1, b1);
    r.a0 = t0; r.a1 = t1; r.a2 = t2;
}
template <class C> HDNI void fp6_inv(Fp6<C>& r, const Fp6<C>& a) {
    Fp2<C> t0, t1, t2, s, d;
    fp2_sqr(t0, a.a0); fp2_mul(s, a.a1, a.a2); fp2_mul_xi(s, s); fp2_sub(t0, t0, s);
    fp2_sqr(t1, a.a2); fp2_mul_xi(t1, t1); fp2_mul(s, a.a0, a.a1); fp2_sub(t1, t1, s);
    fp2_sqr(t2, a.a1); fp2_mul(s, a.a0, a.a2); fp2_sub(t2, t2, s);
    fp2_mul(d, a.a2, t1); fp2_mul(s, a.a1, t2); fp2_add(d, d, s); fp2_mul_xi(d, d);
    fp2_mul(s, a.a0, t0); fp2_add(d, d, s);
    fp2_inv(d, d);
    fp2_mul(r.a0, t0, d);
    fp2_mul(r.a1, t1, d);
    fp2_mul(r.a2, t2, d);
}

// ------------------------------------------------------------------------------------ Fp12
template <class C> HD void fp12_one(Fp12<C>& r) {
    fp2_one(r.c0.a0); fp2_zero(r.c0.a1); fp2_zero(r.c0.a2);
    fp2_zero(r.c1.a0); fp2_zero(r.c1.a1); fp2_zero(r.c1.a2);
}
template <class C> HD bool fp12_is_one(const Fp12<C>& a) {
    Fp2<C> one;
    fp2_one(one);
    return fp2_eq(a.c0.a0, one) && fp2_is_zero(a.c0.a1) && fp2_is_zero(a.c0.a2) && fp2_is_zero(a.c1.a0) &&
           fp2_is_zero(a.c1.a1) && fp2_is_zero(a.c1.a2);
}
template <class C> HDNI void fp12_mul(Fp12<C>& r, const Fp12<C>& a, const Fp12<C>& b) {
    Fp6<C> t0, t1, s, t, u;
    fp6_mul(t0, a.c0, b.c0);
    fp6_mul(t1, a.c1, b.c1);
    fp6_add(s, a.c0, a.c1);
    fp6_add(t, b.c0, b.c1);
    fp6_mul(u, s, t);
    fp6_sub(u, u, t0);
    fp6_sub(r.c1, u, t1);
    fp6_mul_v(t1, t1);
    fp6_add(r.c0, t0, t1);
}
template <class C> HDNI void fp12_sqr(Fp12<C>& r, const Fp12<C>& a) {
    Fp6<C> t, s, u, vt;
    fp6_mul(t, a.c0, a.c1);
    fp6_add(s, a.c0, a.c1);
    fp6_mul_v(u, a.c1);
    fp6_add(u, u, a.c0);
    fp6_mul(s, s, u);
    fp6_mul_v(vt, t);
    fp6_sub(s, s, t);
    fp6_sub(r.c0, s, vt);
    fp6_add(r.c1, t, t);
}
template <class C> HD void fp12_conj(Fp12<C>& r, const Fp12<C>& a) { r.c0 = a.c0; fp6_neg(r.c1, a.c1); }
template <class C> HDNI void fp12_inv(Fp12<C>& r, const Fp12<C>& a) {
    Fp6<C> t0, t1;
    fp6_mul(t0, a.c0, a.c0);
    fp6_mul(t1, a.c1, a.c1);
    fp6_mul_v(t1, t1);
    fp6_sub(t0, t0, t1);
    fp6_inv(t0, t0);
    fp6_mul(r.c0, a.c0, t0);
    fp6_mul(t1, a.c1, t0);
    fp6_neg(r.c1, t1);
}
// coefficient of w^k (k even -> c0.a[k/2], k odd -> c1.a[(k-1)/2])
template <class C> HD Fp2<C>& fp12_coef(Fp12<C>& a, int k) {
    Fp6<C>& h = (k & 1) ? a.c1 : a.c0;
    const int j = k >> 1;
    return j == 0 ? h.a0 : (j == 1 ? h.a1 : h.a2);
}
// a^(p^e), e = 1, 2, 3: coefficient of w^k -> (conj^e c_k) * gamma_e[k]
template <class C> HDNI void fp12_frob(Fp12<C>& r, const Fp12<C>& a, int e) {
    Fp12<C> t = a;
    for (int k = 0; k < 6; k++) {
        Fp2<C>& c = fp12_coef(t, k);
        if (e & 1) fp_neg(c.c1, c.c1);
        if (k == 0) continue;
        if (e == 2) {
            Fp<C> g;
            fp_set(g, C::GAMMA2(k));
            fp2_mul_fp(c, c, g);
        } else {
            Fp2<C> g;
            fp2_set(g, e == 1 ? C::GAMMA1(k) : C::GAMMA3(k));
            fp2_mul(c, c, g);
        }
    }
    r = t;
}
// Granger-Scott squaring in the cyclotomic subgroup
template <class C> HD void fp4_sqr(Fp2<C>& o0, Fp2<C>& o1, const Fp2<C>& a, const Fp2<C>& b) {
    Fp2<C> t0, t1, s;
    fp2_sqr(t0, a);
    fp2_sqr(t1, b);
    fp2_add(s, a, b);
    fp2_sqr(s, s);
    fp2_sub(s, s, t0);
    fp2_sub(o1, s, t1);
    fp2_mul_xi(t1, t1);
    fp2_add(o0, t1, t0);
}
template <class C> HDNI void fp12_cyc_sqr(Fp12<C>& r, const Fp12<C>& f) {
    Fp2<C> z0 = f.c0.a0, z4 = f.c0.a1, z3 = f.c0.a2, z2 = f.c1.a0, z1 = f.c1.a1, z5 = f.c1.a2;
    Fp2<C> t0, t1, t2, t3, s;
    fp4_sqr(t0, t1, z0, z1);
    fp2_sub(s, t0, z0); fp2_dbl(s, s); fp2_add(z0, s, t0);
    fp2_add(s, t1, z1); fp2_dbl(s, s); fp2_add(z1, s, t1);
    fp4_sqr(t0, t1, z2, z3);
    fp4_sqr(t2, t3, z4, z5);
    fp2_sub(s, t0, z4); fp2_dbl(s, s); fp2_add(z4, s, t0);
    fp2_add(s, t1, z5); fp2_dbl(s, s); fp2_add(z5, s, t1);
    fp2_mul_xi(t0, t3);
    fp2_add(s, t0, z2); fp2_dbl(s, s); fp2_add(z2, s, t0);
    fp2_sub(s, t2, z3); fp2_dbl(s, s); fp2_add(z3, s, t2);
    r.c0.a0 = z0; r.c0.a1 = z4; r.c0.a2 = z3;
    r.c1.a0 = z2; r.c1.a1 = z1; r.c1.a2 = z5;
}
// a^e for unitary a, e < 2^128 given as (hi, lo)
template <class C> HDNI void fp12_cyc_pow(Fp12<C>& r, const Fp12<C>& a, unsigned long long hi, unsigned long long lo) {
    Fp12<C> acc, base = a;
    fp12_one(acc);
    bool started = false;
    for (int i = 127; i >= 0; i--) {
        if (started) fp12_cyc_sqr(acc, acc);
        const unsigned long long w = i >= 64 ? hi : lo;
        if ((w >> (i & 63)) & 1) {
            if (started) fp12_mul(acc, acc, base);
            else { acc = base; started = true; }
        }
    }
    r = acc;
}

// GT wire layout (12 x FP_BYTES big-endian): w-powers 5,3,1,4,2,0, each (im, re)
template <class C> HD void fp12_to_be(uint8_t* out, const Fp12<C>& a) {
    Fp12<C> t = a;
    const int order[6] = {5, 3, 1, 4, 2, 0};
    for (int i = 0; i < 6; i++) {
        const Fp2<C>& c = fp12_coef(t, order[i]);
        fp_to_be<C>(out + (2 * i) * C::FP_BYTES, c.c1);
        fp_to_be<C>(out + (2 * i + 1) * C::FP_BYTES, c.c0);
    }
}
template <class C> HD void fp12_from_be(Fp12<C>& a, const uint8_t* in) {
    const int order[6] = {5, 3, 1, 4, 2, 0};
    for (int i = 0; i < 6; i++) {
        Fp2<C>& c = fp12_coef(a, order[i]);
        fp_from_be<C>(c.c1, in + (2 * i) * C::FP_BYTES);
        fp_from_be<C>(c.c0, in + (2 * i + 1) * C::FP_BYTES);
    }
}

}  // namespace bgls
