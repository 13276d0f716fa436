// Optimal-ate Miller loop, final exponentiation and the G1/G2 group law, generic in the curve
// traits.  Replaces the arithmetic behind the reference's
//   CurveSystem.Pair / PairingProduct   curves/curve.go:46-48,125-170; altbn128.go:130-145; bls12_381.go:228-240
//   Point.Add / AggregatePoints         curves/curve.go:73-121; altbn128.go:59-66,181-188; bls12_381.go:33-41,94-102
//   Point.Mul / ScalePoints             curves/curve.go:190-214; altbn128.go:107-121,235-249
// Wire formats are the reference's uncompressed affine big-endian records (altbn128.go:149-158,
// bls12_381.go:147-158): G1 = x||y, G2 = x_im||x_re||y_im||y_re; infinity = all-zero record.
#pragma once
#include "field.cuh"

namespace bgls {

template <class C> struct G1Aff { Fp<C> x, y; bool inf; };
template <class C> struct G2Aff { Fp2<C> x, y; bool inf; };
template <class C> struct G2Proj { Fp2<C> X, Y, Z; };

// ---------------------------------------------------------------- wire format
HD bool bytes_all_zero(const uint8_t* p, int n) {
    uint32_t x = 0;
    for (int i = 0; i < n; i++) x |= p[i];
    return x == 0;
}
template <class C> HD void g1_load(G1Aff<C>& P, const uint8_t* in) {
    P.inf = bytes_all_zero(in, 2 * C::FP_BYTES) || (!C::IS_BN && (in[0] & 0x40));
    if (P.inf) { fp_zero(P.x); fp_zero(P.y); return; }
    fp_from_be<C>(P.x, in);
    fp_from_be<C>(P.y, in + C::FP_BYTES);
}
template <class C> HD void g2_load(G2Aff<C>& Q, const uint8_t* in) {
    Q.inf = bytes_all_zero(in, 4 * C::FP_BYTES) || (!C::IS_BN && (in[0] & 0x40));
    if (Q.inf) { fp2_zero(Q.x); fp2_zero(Q.y); return; }
    fp_from_be<C>(Q.x.c1, in);
    fp_from_be<C>(Q.x.c0, in + C::FP_BYTES);
    fp_from_be<C>(Q.y.c1, in + 2 * C::FP_BYTES);
    fp_from_be<C>(Q.y.c0, in + 3 * C::FP_BYTES);
}

// ---------------------------------------------------------------- sparse line multiplication
// altbn128 (D-type twist): line = l0 + l1 w + l3 w^3      -> (c0,c1) = ((l0,0,0), (l1,l3,0))
// bls12-381 (M-type twist): line = l0 + l2 w^2 + l3 w^3   -> (c0,c1) = ((l0,l2,0), (0,l3,0))
// Arguments are always passed as (yP-term, xP-term, constant term).
template <class C> HDNI void fp12_mul_line(Fp12<C>& f, const Fp2<C>& ly, const Fp2<C>& lx, const Fp2<C>& lc) {
    Fp6<C> t0, t1, s, u;
    Fp2<C> e;
    if (C::IS_BN) {
        fp6_mul_by_0(t0, f.c0, ly);
        fp6_mul_by_01(t1, f.c1, lx, lc);
        fp6_add(s, f.c0, f.c1);
        fp2_add(e, ly, lx);
        fp6_mul_by_01(u, s, e, lc);
    } else {
        fp6_mul_by_01(t0, f.c0, lc, lx);
        fp6_mul_by_1(t1, f.c1, ly);
        fp6_add(s, f.c0, f.c1);
        fp2_add(e, lx, ly);
        fp6_mul_by_01(u, s, lc, e);
    }
    fp6_sub(u, u, t0);
    fp6_sub(f.c1, u, t1);
    fp6_mul_v(t1, t1);
    fp6_add(f.c0, t0, t1);
}

// ---------------------------------------------------------------- Miller loop steps
// T <- 2T and the tangent line at T evaluated at P (homogeneous projective, a = 0):
//   H = 2YZ, B = Y^2, E = 3b'Z^2 :  line = H yP  +  (-3 X^2 xP) [w | w^2]  +  (B - E) [w^3 | 1]
template <class C> HDNI void dbl_step(Fp12<C>& f, G2Proj<C>& T, const G1Aff<C>& P) {
    Fp2<C> A, B, Cc, E, F, G, H, X2, t, ly, lx, lc, b3;
    Fp<C> half;
    fp_set(half, C::HALF());
    fp2_set(b3, C::B2X3());
    fp2_mul(A, T.X, T.Y);
    fp2_mul_fp(A, A, half);
    fp2_sqr(B, T.Y);
    fp2_sqr(Cc, T.Z);
    fp2_mul(E, Cc, b3);
    fp2_dbl(F, E);
    fp2_add(F, F, E);
    fp2_add(G, B, F);
    fp2_mul_fp(G, G, half);
    fp2_add(H, T.Y, T.Z);
    fp2_sqr(H, H);
    fp2_sub(H, H, B);
    fp2_sub(H, H, Cc);
    fp2_sqr(X2, T.X);
    fp2_mul_fp(ly, H, P.y);
    fp2_dbl(t, X2);
    fp2_add(t, t, X2);
    fp2_mul_fp(lx, t, P.x);
    fp2_neg(lx, lx);
    fp2_sub(lc, B, E);
    fp2_sub(t, B, F);
    fp2_mul(T.X, A, t);
    fp2_sqr(t, G);
    fp2_sqr(Cc, E);
    fp2_dbl(A, Cc);
    fp2_add(A, A, Cc);
    fp2_sub(T.Y, t, A);
    fp2_mul(T.Z, B, H);
    fp12_mul_line(f, ly, lx, lc);
}
// T <- T + Q (Q affine) and the chord through T, Q evaluated at P:
//   theta = Y - yQ Z, lam = X - xQ Z : line = lam yP + (-theta xP)[..] + (theta xQ - lam yQ)[..]
template <class C> HDNI void add_step(Fp12<C>& f, G2Proj<C>& T, const G2Aff<C>& Q, const G1Aff<C>& P) {
    Fp2<C> th, la, Cc, D, E, F, G, H, t, ly, lx, lc;
    fp2_mul(t, Q.y, T.Z);
    fp2_sub(th, T.Y, t);
    fp2_mul(t, Q.x, T.Z);
    fp2_sub(la, T.X, t);
    fp2_mul_fp(ly, la, P.y);
    fp2_mul_fp(lx, th, P.x);
    fp2_neg(lx, lx);
    fp2_mul(lc, th, Q.x);
    fp2_mul(t, la, Q.y);
    fp2_sub(lc, lc, t);
    fp2_sqr(Cc, th);
    fp2_sqr(D, la);
    fp2_mul(E, la, D);
    fp2_mul(F, T.Z, Cc);
    fp2_mul(G, T.X, D);
    fp2_add(H, E, F);
    fp2_sub(H, H, G);
    fp2_sub(H, H, G);
    fp2_mul(T.X, la, H);
    fp2_sub(t, G, H);
    fp2_mul(t, th, t);
    fp2_mul(G, E, T.Y);
    fp2_sub(T.Y, t, G);
    fp2_mul(T.Z, T.Z, E);
    fp12_mul_line(f, ly, lx, lc);
}

// f_{lambda,Q}(P); 1 when either point is infinity (the reference *defines* the GT identity as
// Pair(G1, inf) / Pair(inf, G2): curves/altbn128.go:478, curves/bls12_381.go:341)
template <class C> HDNI void miller_loop(Fp12<C>& f, const G1Aff<C>& P, const G2Aff<C>& Q) {
    fp12_one(f);
    if (P.inf || Q.inf) return;
    G2Proj<C> T;
    T.X = Q.x;
    T.Y = Q.y;
    fp2_one(T.Z);
    for (int i = C::LOOP_TOP - 1; i >= 0; i--) {
        if (i != C::LOOP_TOP - 1) fp12_sqr(f, f);  // f == 1 on the first pass
        dbl_step(f, T, P);
        if ((C::LOOP_LO >> i) & 1) add_step(f, T, Q, P);
    }
    if (C::IS_BN) {
        G2Aff<C> Q1, Q2;
        Fp2<C> g2, g3, t;
        Q1.inf = Q2.inf = false;
        fp2_set(g2, C::GAMMA1(2));
        fp2_set(g3, C::GAMMA1(3));
        fp2_conj(t, Q.x); fp2_mul(Q1.x, t, g2);
        fp2_conj(t, Q.y); fp2_mul(Q1.y, t, g3);
        fp2_conj(t, Q1.x); fp2_mul(Q2.x, t, g2);
        fp2_conj(t, Q1.y); fp2_mul(Q2.y, t, g3);
        fp2_neg(Q2.y, Q2.y);
        add_step(f, T, Q1, P);
        add_step(f, T, Q2, P);
    } else {
        fp12_conj(f, f);  // x < 0
    }
}

// Product of the Miller functions of k <= K pairs with ONE shared accumulator: prod_j f_{lambda,Q_j}(P_j).
// Every iteration squares f once and multiplies the k lines into it, instead of k squarings (the squaring is
// 36 of the 102 Fp multiplications of a doubling iteration).  Pairs with a point at infinity contribute 1.
template <class C, int K> HDNI void miller_loop_shared(Fp12<C>& f, const G1Aff<C>* P, const G2Aff<C>* Q, int k) {
    fp12_one(f);
    G2Proj<C> T[K];
    bool any = false;
    for (int j = 0; j < k; j++) {
        T[j].X = Q[j].x;
        T[j].Y = Q[j].y;
        fp2_one(T[j].Z);
        any = any || !(P[j].inf || Q[j].inf);
    }
    if (!any) return;
    for (int i = C::LOOP_TOP - 1; i >= 0; i--) {
        if (i != C::LOOP_TOP - 1) fp12_sqr(f, f);  // f == 1 on the first pass
        const bool bit = (C::LOOP_LO >> i) & 1;
        for (int j = 0; j < k; j++) {
            if (P[j].inf || Q[j].inf) continue;
            dbl_step(f, T[j], P[j]);
            if (bit) add_step(f, T[j], Q[j], P[j]);
        }
    }
    if (C::IS_BN) {
        Fp2<C> g2, g3, t;
        fp2_set(g2, C::GAMMA1(2));
        fp2_set(g3, C::GAMMA1(3));
        for (int j = 0; j < k; j++) {
            if (P[j].inf || Q[j].inf) continue;
            G2Aff<C> Q1, Q2;
            Q1.inf = Q2.inf = false;
            fp2_conj(t, Q[j].x); fp2_mul(Q1.x, t, g2);
            fp2_conj(t, Q[j].y); fp2_mul(Q1.y, t, g3);
            fp2_conj(t, Q1.x); fp2_mul(Q2.x, t, g2);
            fp2_conj(t, Q1.y); fp2_mul(Q2.y, t, g3);
            fp2_neg(Q2.y, Q2.y);
            add_step(f, T[j], Q1, P[j]);
            add_step(f, T[j], Q2, P[j]);
        }
    } else {
        fp12_conj(f, f);  // x < 0
    }
}

// ---------------------------------------------------------------- final exponentiation
template <class C> HDNI void final_exp(Fp12<C>& r, const Fp12<C>& f) {
    Fp12<C> m, t0, t1;
    // easy part: f^((p^6-1)(p^2+1))
    fp12_conj(t0, f);
    fp12_inv(t1, f);
    fp12_mul(t0, t0, t1);
    fp12_frob(t1, t0, 2);
    fp12_mul(m, t1, t0);
    if constexpr (C::IS_BN) {
        // exact (p^4-p^2+1)/r power, Devegili-Scott-Dahab addition chain with three u-powers
        Fp12<C> fu, fu2, fu3, y0, y1, y2, y3, y4, y5, y6;
        fp12_cyc_pow(fu, m, 0, C::U);
        fp12_cyc_pow(fu2, fu, 0, C::U);
        fp12_cyc_pow(fu3, fu2, 0, C::U);
        fp12_frob(y3, fu, 1);
        fp12_conj(y3, y3);
        fp12_frob(t0, fu2, 1);      // fu2p
        fp12_mul(y4, fu, t0);
        fp12_conj(y4, y4);
        fp12_frob(t1, fu3, 1);      // fu3p
        fp12_mul(y6, fu3, t1);
        fp12_conj(y6, y6);
        fp12_frob(y2, fu2, 2);
        fp12_frob(y0, m, 1);
        fp12_frob(t0, m, 2);
        fp12_mul(y0, y0, t0);
        fp12_frob(t0, m, 3);
        fp12_mul(y0, y0, t0);
        fp12_conj(y1, m);
        fp12_conj(y5, fu2);
        fp12_cyc_sqr(t0, y6);
        fp12_mul(t0, t0, y4);
        fp12_mul(t0, t0, y5);
        fp12_mul(t1, y3, y5);
        fp12_mul(t1, t1, t0);
        fp12_mul(t0, t0, y2);
        fp12_cyc_sqr(t1, t1);
        fp12_mul(t1, t1, t0);
        fp12_cyc_sqr(t1, t1);
        fp12_mul(t0, t1, y1);
        fp12_mul(t1, t1, y0);
        fp12_cyc_sqr(t0, t0);
        fp12_mul(r, t0, t1);
    } else {
        // hard = ((x-1)^2/3)(x+p)(x^2+p^2-1) + 1 (exact), x = -|x|
        Fp12<C> y0, y1, y2;
        fp12_cyc_pow(y0, m, C::C_HI, C::C_LO);
        fp12_cyc_pow(t0, y0, 0, C::U);
        fp12_conj(t0, t0);
        fp12_frob(t1, y0, 1);
        fp12_mul(y1, t0, t1);
        fp12_cyc_pow(t0, y1, 0, C::U);
        fp12_cyc_pow(t0, t0, 0, C::U);
        fp12_frob(t1, y1, 2);
        fp12_mul(y2, t0, t1);
        fp12_conj(t0, y1);
        fp12_mul(y2, y2, t0);
        fp12_mul(r, y2, m);
    }
}

// ---------------------------------------------------------------- group law (Jacobian), F = Fp<C> or Fp2<C>
template <class C> HD void fe_add(Fp<C>& r, const Fp<C>& a, const Fp<C>& b) { fp_add(r, a, b); }
template <class C> HD void fe_sub(Fp<C>& r, const Fp<C>& a, const Fp<C>& b) { fp_sub(r, a, b); }
template <class C> HDNI void fe_mul(Fp<C>& r, const Fp<C>& a, const Fp<C>& b) { fp_mul(r, a, b); }
template <class C> HDNI void fe_sqr(Fp<C>& r, const Fp<C>& a) { fp_sqr(r, a); }
template <class C> HD void fe_inv(Fp<C>& r, const Fp<C>& a) { fp_inv(r, a); }
template <class C> HD bool fe_is_zero(const Fp<C>& a) { return fp_is_zero(a); }
template <class C> HD bool fe_eq(const Fp<C>& a, const Fp<C>& b) { return fp_eq(a, b); }
template <class C> HD void fe_one(Fp<C>& r) { fp_set(r, C::R1()); }
template <class C> HD void fe_add(Fp2<C>& r, const Fp2<C>& a, const Fp2<C>& b) { fp2_add(r, a, b); }
template <class C> HD void fe_sub(Fp2<C>& r, const Fp2<C>& a, const Fp2<C>& b) { fp2_sub(r, a, b); }
template <class C> HD void fe_mul(Fp2<C>& r, const Fp2<C>& a, const Fp2<C>& b) { fp2_mul(r, a, b); }
template <class C> HD void fe_sqr(Fp2<C>& r, const Fp2<C>& a) { fp2_sqr(r, a); }
template <class C> HD void fe_inv(Fp2<C>& r, const Fp2<C>& a) { fp2_inv(r, a); }
template <class C> HD bool fe_is_zero(const Fp2<C>& a) { return fp2_is_zero(a); }
template <class C> HD bool fe_eq(const Fp2<C>& a, const Fp2<C>& b) { return fp2_eq(a, b); }
template <class C> HD void fe_one(Fp2<C>& r) { fp2_one(r); }

template <class F> struct Jac { F X, Y, Z; bool inf; };

template <class F> HDNI void jac_dbl(Jac<F>& r, const Jac<F>& p) {
    if (p.inf || fe_is_zero(p.Y)) { r.inf = true; return; }
    F A, B, Cc, D, E, Ff, t;
    Jac<F> o;
    o.inf = false;
    fe_sqr(A, p.X);
    fe_sqr(B, p.Y);
    fe_sqr(Cc, B);
    fe_add(t, p.X, B);
    fe_sqr(t, t);
    fe_sub(t, t, A);
    fe_sub(t, t, Cc);
    fe_add(D, t, t);
    fe_add(E, A, A);
    fe_add(E, E, A);
    fe_sqr(Ff, E);
    fe_sub(t, Ff, D);
    fe_sub(o.X, t, D);
    fe_mul(t, p.Y, p.Z);
    fe_add(o.Z, t, t);
    fe_sub(t, D, o.X);
    fe_mul(t, E, t);
    fe_add(Cc, Cc, Cc);
    fe_add(Cc, Cc, Cc);
    fe_add(Cc, Cc, Cc);
    fe_sub(o.Y, t, Cc);
    r = o;
}
template <class F> HDNI void jac_add(Jac<F>& r, const Jac<F>& p, const Jac<F>& q) {
    if (p.inf) { r = q; return; }
    if (q.inf) { r = p; return; }
    F Z1Z1, Z2Z2, U1, U2, S1, S2, H, R, t, HH, HHH, V;
    F one;
    fe_one(one);
    const bool q_affine = fe_eq(q.Z, one);   // mixed addition (the addend of jac_mul, wire points): 4 products fewer
    fe_sqr(Z1Z1, p.Z);
    if (q_affine) { U1 = p.X; S1 = p.Y; }
    else {
        fe_sqr(Z2Z2, q.Z);
        fe_mul(U1, p.X, Z2Z2);
        fe_mul(t, q.Z, Z2Z2);
        fe_mul(S1, p.Y, t);
    }
    fe_mul(U2, q.X, Z1Z1);
    fe_mul(t, p.Z, Z1Z1);
    fe_mul(S2, q.Y, t);
    if (fe_eq(U1, U2)) {
        if (fe_eq(S1, S2)) { jac_dbl(r, p); return; }
        r.inf = true;
        return;
    }
    Jac<F> o;
    o.inf = false;
    fe_sub(H, U2, U1);
    fe_sub(R, S2, S1);
    fe_sqr(HH, H);
    fe_mul(HHH, HH, H);
    fe_mul(V, U1, HH);
    fe_sqr(t, R);
    fe_sub(t, t, HHH);
    fe_sub(t, t, V);
    fe_sub(o.X, t, V);
    fe_sub(t, V, o.X);
    fe_mul(t, R, t);
    fe_mul(S1, S1, HHH);
    fe_sub(o.Y, t, S1);
    if (q_affine) fe_mul(o.Z, p.Z, H);
    else {
        fe_mul(t, p.Z, q.Z);
        fe_mul(o.Z, t, H);
    }
    r = o;
}
// scalar: 32 bytes big-endian
template <class F> HDNI void jac_mul(Jac<F>& r, const Jac<F>& p, const uint8_t* scalar32) {
    Jac<F> acc;
    acc.inf = true;
    for (int i = 0; i < 256; i++) {
        jac_dbl(acc, acc);
        if ((scalar32[i >> 3] >> (7 - (i & 7))) & 1) jac_add(acc, acc, p);
    }
    r = acc;
}
template <class F> HD void jac_to_affine(F& x, F& y, bool& inf, const Jac<F>& p) {
    inf = p.inf || fe_is_zero(p.Z);
    if (inf) return;
    F zi, zi2, zi3;
    fe_inv(zi, p.Z);
    fe_sqr(zi2, zi);
    fe_mul(zi3, zi2, zi);
    fe_mul(x, p.X, zi2);
    fe_mul(y, p.Y, zi3);
}

template <class C> HD void jac_load(Jac<Fp<C>>& r, const uint8_t* in) {
    G1Aff<C> a;
    g1_load<C>(a, in);
    r.X = a.x; r.Y = a.y; r.inf = a.inf;
    fe_one(r.Z);
}
template <class C> HD void jac_load(Jac<Fp2<C>>& r, const uint8_t* in) {
    G2Aff<C> a;
    g2_load<C>(a, in);
    r.X = a.x; r.Y = a.y; r.inf = a.inf;
    fe_one(r.Z);
}
template <class C> HD void jac_store(uint8_t* out, const Jac<Fp<C>>& p) {
    Fp<C> x, y;
    bool inf;
    jac_to_affine(x, y, inf, p);
    if (inf) { for (int i = 0; i < 2 * C::FP_BYTES; i++) out[i] = 0; return; }
    fp_to_be<C>(out, x);
    fp_to_be<C>(out + C::FP_BYTES, y);
}
template <class C> HD void jac_store(uint8_t* out, const Jac<Fp2<C>>& p) {
    Fp2<C> x, y;
    bool inf;
    jac_to_affine(x, y, inf, p);
    if (inf) { for (int i = 0; i < 4 * C::FP_BYTES; i++) out[i] = 0; return; }
    fp_to_be<C>(out, x.c1);
    fp_to_be<C>(out + C::FP_BYTES, x.c0);
    fp_to_be<C>(out + 2 * C::FP_BYTES, y.c1);
    fp_to_be<C>(out + 3 * C::FP_BYTES, y.c0);
}

}  // namespace bgls
