// Compressed wire formats on the device: Point.Marshal and CurveSystem.UnmarshalG1/G2 on
// compressed input.
//   altbn128  -- the reference's own codec, curves/altbn128.go:81-89 (G1 Marshal), :203-221 (G2 Marshal),
//                :296-376 (Unmarshal), square roots curves/hash.go:178-223:
//                  G1: x (32 B big-endian), bit 7 of byte 0 set iff 2y > q;           infinity = zeros
//                  G2: x_im || x_re (64 B), bit 7 of x_im[0] iff 2 y_im > q, bit 7 of x_re[0] iff 2 y_re > q
//   bls12-381 -- the reference delegates to dis2/bls12 (curves/bls12_381.go:57-63,118-124,242-264), absent from the
//                reference tree; restated as the zcash serialisation that library implements [parity unpinned]:
//                  G1: x (48 B), byte 0: 0x80 compressed, 0x40 infinity, 0x20 y is the lexicographically larger root
//                  G2: x_c1 || x_c0 (96 B), same flags, "larger" compares (y_c1, y_c0)
// A decompressed record is the uncompressed affine layout the rest of the engine uses; ok = 0 marks an input
// that is malformed, not reduced, or not the abscissa of a curve point (the reference returns (nil, false)).
#pragma once
#include "hash.cuh"   // fp_sqrt_candidate, fp_is_quad_res, fp_parity
#include "pairing.cuh"

namespace bgls {

DEVCONST uint8_t BN254_ORDER_BE[32] = {0x30, 0x64, 0x4e, 0x72, 0xe1, 0x31, 0xa0, 0x29, 0xb8, 0x50, 0x45, 0xb6, 0x81, 0x81, 0x58, 0x5d,
                                       0x28, 0x33, 0xe8, 0x48, 0x79, 0xb9, 0x70, 0x91, 0x43, 0xe1, 0xf5, 0x93, 0xf0, 0x00, 0x00, 0x01};
DEVCONST uint8_t BLS381_ORDER_BE[32] = {0x73, 0xed, 0xa7, 0x53, 0x29, 0x9d, 0x7d, 0x48, 0x33, 0x39, 0xd8, 0x08, 0x09, 0xa1, 0xd8, 0x05,
                                        0x53, 0xbd, 0xa4, 0x02, 0xff, 0xfe, 0x5b, 0xfe, 0xff, 0xff, 0xff, 0xff, 0x00, 0x00, 0x00, 0x01};
template <class C> HD const uint8_t* order_be() { return C::IS_BN ? BN254_ORDER_BE : BLS381_ORDER_BE; }

// big-endian field bytes (flag bits already cleared) strictly below p ?
template <class C> HD bool be_below_p(const uint8_t* be) {
    for (int i = C::N - 1; i >= 0; i--) {
        const uint8_t* q = be + 4 * (C::N - 1 - i);
        const uint32_t w = ((uint32_t)q[0] << 24) | ((uint32_t)q[1] << 16) | ((uint32_t)q[2] << 8) | q[3];
        if (w != C::p(i)) return w < C::p(i);
    }
    return false;
}

// square root in Fp2 = Fp[i]/(i^2+1) by the complex method (calcComplexQuadRes, curves/hash.go:196-223);
// returns true iff r*r == a
template <class C> HDNI bool fp2_sqrt(Fp2<C>& r, const Fp2<C>& a) {
    Fp<C> half, lam, delta, t;
    fp_set(half, C::HALF());
    if (fp_is_zero(a.c1)) {
        if (fp_is_quad_res(a.c0)) {
            fp_sqrt_candidate(r.c0, a.c0);
            fp_zero(r.c1);
        } else {
            fp_neg(t, a.c0);
            fp_zero(r.c0);
            fp_sqrt_candidate(r.c1, t);
        }
    } else {
        fp_sqr(lam, a.c0);
        fp_sqr(t, a.c1);
        fp_add(lam, lam, t);
        fp_sqrt_candidate(lam, lam);
        fp_add(delta, a.c0, lam);
        fp_mul(delta, delta, half);
        if (!fp_is_quad_res(delta)) {
            fp_sub(delta, a.c0, lam);
            fp_mul(delta, delta, half);
        }
        fp_sqrt_candidate(r.c0, delta);
        fp_inv(t, r.c0);
        fp_mul(t, t, half);
        fp_mul(r.c1, t, a.c1);
    }
    Fp2<C> chk;
    fp2_sqr(chk, r);
    return fp2_eq(chk, a);
}

template <class C> HD void g1_y2(Fp<C>& y2, const Fp<C>& x) {
    Fp<C> b;
    fp_set(b, C::B1());
    fp_sqr(y2, x);
    fp_mul(y2, y2, x);
    fp_add(y2, y2, b);
}
template <class C> HD void g2_y2(Fp2<C>& y2, const Fp2<C>& x) {
    Fp2<C> b;
    fp2_set(b, C::B2());
    fp2_sqr(y2, x);
    fp2_mul(y2, y2, x);
    fp2_add(y2, y2, b);
}
// zcash ordering: y is the "larger" root (compare c1 first, then c0)
template <class C> HD bool fp2_is_larger(const Fp2<C>& y) { return fp_is_zero(y.c1) ? fp_parity(y.c0) : fp_parity(y.c1); }

// ---------------------------------------------------------------- compress (uncompressed record -> compressed)
template <class C> HDNI void compress_g1(uint8_t* out, const uint8_t* rec) {
    constexpr int F = C::FP_BYTES;
    G1Aff<C> P;
    g1_load<C>(P, rec);
    if (P.inf) {
        for (int i = 0; i < F; i++) out[i] = 0;
        if (!C::IS_BN) out[0] = 0xC0;
        return;
    }
    for (int i = 0; i < F; i++) out[i] = rec[i];
    const bool big = fp_parity(P.y);
    if (C::IS_BN) { if (big) out[0] |= 0x80; }
    else out[0] |= 0x80 | (big ? 0x20 : 0);
}
template <class C> HDNI void compress_g2(uint8_t* out, const uint8_t* rec) {
    constexpr int F = C::FP_BYTES;
    G2Aff<C> Q;
    g2_load<C>(Q, rec);
    if (Q.inf) {
        for (int i = 0; i < 2 * F; i++) out[i] = 0;
        if (!C::IS_BN) out[0] = 0xC0;
        return;
    }
    for (int i = 0; i < 2 * F; i++) out[i] = rec[i];   // x_im || x_re
    if (C::IS_BN) {
        if (fp_parity(Q.y.c1)) out[0] |= 0x80;
        if (fp_parity(Q.y.c0)) out[F] |= 0x80;
    } else {
        out[0] |= 0x80 | (fp2_is_larger(Q.y) ? 0x20 : 0);
    }
}

// ---------------------------------------------------------------- decompress; returns ok
// check_subgroup: additionally require r*P = infinity (the reference's bls12 Unmarshal calls Check(),
// curves/bls12_381.go:248,260; its altbn128 path does not)
template <class C> HDNI bool decompress_g1(uint8_t* rec, const uint8_t* in, bool check_subgroup) {
    constexpr int F = C::FP_BYTES;
    uint8_t xb[F];
    for (int i = 0; i < F; i++) xb[i] = in[i];
    for (int i = 0; i < 2 * F; i++) rec[i] = 0;
    bool sgn, inf = false;
    if (C::IS_BN) {
        sgn = (xb[0] & 0x80) != 0;
        xb[0] &= 0x7F;
    } else {
        if (!(xb[0] & 0x80)) return false;
        inf = (xb[0] & 0x40) != 0;
        sgn = (xb[0] & 0x20) != 0;
        xb[0] &= 0x1F;
    }
    const bool zero = bytes_all_zero(xb, F);
    if (C::IS_BN) { if (zero) return true; }
    else if (inf) return zero && !sgn;
    if (!be_below_p<C>(xb)) return false;
    Fp<C> x, y, y2, chk;
    fp_from_be<C>(x, xb);
    g1_y2<C>(y2, x);
    fp_sqrt_candidate(y, y2);
    fp_sqr(chk, y);
    if (!fp_eq(chk, y2)) return false;
    if (sgn != fp_parity(y)) fp_neg(y, y);
    if (check_subgroup) {
        Jac<Fp<C>> p, q;
        p.X = x; p.Y = y; p.inf = false;
        fe_one(p.Z);
        jac_mul(q, p, order_be<C>());
        if (!(q.inf || fe_is_zero(q.Z))) return false;
    }
    for (int i = 0; i < F; i++) rec[i] = xb[i];
    fp_to_be<C>(rec + F, y);
    return true;
}
template <class C> HDNI bool decompress_g2(uint8_t* rec, const uint8_t* in, bool check_subgroup) {
    constexpr int F = C::FP_BYTES;
    uint8_t xb[2 * F];
    for (int i = 0; i < 2 * F; i++) xb[i] = in[i];
    for (int i = 0; i < 4 * F; i++) rec[i] = 0;
    bool s_im, s_re = false, inf = false;
    if (C::IS_BN) {
        s_im = (xb[0] & 0x80) != 0;
        s_re = (xb[F] & 0x80) != 0;
        xb[0] &= 0x7F;
        xb[F] &= 0x7F;
    } else {
        if (!(xb[0] & 0x80)) return false;
        inf = (xb[0] & 0x40) != 0;
        s_im = (xb[0] & 0x20) != 0;
        xb[0] &= 0x1F;
    }
    const bool zero = bytes_all_zero(xb, 2 * F);
    if (C::IS_BN) { if (zero) return true; }
    else if (inf) return zero && !s_im;
    if (!be_below_p<C>(xb) || !be_below_p<C>(xb + F)) return false;
    Fp2<C> x, y, y2;
    fp_from_be<C>(x.c1, xb);
    fp_from_be<C>(x.c0, xb + F);
    g2_y2<C>(y2, x);
    const bool is_root = fp2_sqrt(y, y2);
    if (C::IS_BN) {
        // the reference fixes the two components independently (curves/altbn128.go:355-370) and then lets the
        // curve library reject what is not on the curve
        if (s_im != fp_parity(y.c1)) fp_neg(y.c1, y.c1);
        if (s_re != fp_parity(y.c0)) fp_neg(y.c0, y.c0);
        Fp2<C> chk;
        fp2_sqr(chk, y);
        if (!fp2_eq(chk, y2)) return false;
    } else {
        if (!is_root) return false;
        if (s_im != fp2_is_larger(y)) fp2_neg(y, y);
    }
    if (check_subgroup) {
        Jac<Fp2<C>> p, q;
        p.X = x; p.Y = y; p.inf = false;
        fe_one(p.Z);
        jac_mul(q, p, order_be<C>());
        if (!(q.inf || fe_is_zero(q.Z))) return false;
    }
    for (int i = 0; i < 2 * F; i++) rec[i] = xb[i];
    fp_to_be<C>(rec + 2 * F, y.c1);
    fp_to_be<C>(rec + 3 * F, y.c0);
    return true;
}

// ---------------------------------------------------------------- validation of uncompressed records
// What the reference enforces when a Point is built from coordinates or bytes (the engine's kernels assume it):
//   altbn128  bn256.G1/G2.Unmarshal (curves/altbn128.go:42-57,160-179,296-376): coordinates < q, on the curve, and for G2
//             membership in the order-r subgroup (the library's twist-point check multiplies by the order)
//   bls12-381 Unmarshal + Check() (curves/bls12_381.go:242-264), MakeG*Point(check = true) (:197-226): on the curve and in
//             the order-r subgroup
// subgroup = false checks range and curve equation only.  Infinity records are valid.
template <class C> HDNI bool validate_g1(const uint8_t* rec, bool subgroup) {
    constexpr int F = C::FP_BYTES;
    if (!C::IS_BN && (rec[0] & 0x40)) {   // bls12 infinity flag: everything else must be zero
        uint32_t x = rec[0] & 0xBF;
        for (int i = 1; i < 2 * F; i++) x |= rec[i];
        return x == 0;
    }
    if (bytes_all_zero(rec, 2 * F)) return true;
    if (!be_below_p<C>(rec) || !be_below_p<C>(rec + F)) return false;
    Fp<C> x, y, y2, chk;
    fp_from_be<C>(x, rec);
    fp_from_be<C>(y, rec + F);
    g1_y2<C>(y2, x);
    fp_sqr(chk, y);
    if (!fp_eq(chk, y2)) return false;
    if (subgroup && !C::IS_BN) {   // altbn128 G1 has cofactor 1
        Jac<Fp<C>> p, q;
        p.X = x; p.Y = y; p.inf = false;
        fe_one(p.Z);
        jac_mul(q, p, order_be<C>());
        if (!(q.inf || fe_is_zero(q.Z))) return false;
    }
    return true;
}
template <class C> HDNI bool validate_g2(const uint8_t* rec, bool subgroup) {
    constexpr int F = C::FP_BYTES;
    if (!C::IS_BN && (rec[0] & 0x40)) {
        uint32_t x = rec[0] & 0xBF;
        for (int i = 1; i < 4 * F; i++) x |= rec[i];
        return x == 0;
    }
    if (bytes_all_zero(rec, 4 * F)) return true;
    for (int k = 0; k < 4; k++)
        if (!be_below_p<C>(rec + k * F)) return false;
    Fp2<C> x, y, y2, chk;
    fp_from_be<C>(x.c1, rec);
    fp_from_be<C>(x.c0, rec + F);
    fp_from_be<C>(y.c1, rec + 2 * F);
    fp_from_be<C>(y.c0, rec + 3 * F);
    g2_y2<C>(y2, x);
    fp2_sqr(chk, y);
    if (!fp2_eq(chk, y2)) return false;
    if (subgroup) {
        Jac<Fp2<C>> p, q;
        p.X = x; p.Y = y; p.inf = false;
        fe_one(p.Z);
        jac_mul(q, p, order_be<C>());
        if (!(q.inf || fe_is_zero(q.Z))) return false;
    }
    return true;
}

// ---------------------------------------------------------------- GT exponentiation (PointT.Mul)
// out = a^e for a in GT (unitary: the cyclotomic squaring applies; the inverse is the conjugate), e = 32-byte big-endian
// magnitude with a sign -- curves/altbn128.go:290-294 (ScalarMult), curves/bls12_381.go:186-195.
template <class C> HDNI void gt_pow(uint8_t* out, const uint8_t* a_be, const uint8_t* e32, bool negative) {
    Fp12<C> a, acc;
    fp12_from_be<C>(a, a_be);
    if (negative) fp12_conj(a, a);
    fp12_one(acc);
    bool started = false;
    for (int i = 0; i < 256; i++) {
        if (started) fp12_cyc_sqr(acc, acc);
        if ((e32[i >> 3] >> (7 - (i & 7))) & 1) {
            if (started) fp12_mul(acc, acc, a);
            else { acc = a; started = true; }
        }
    }
    fp12_to_be<C>(out, acc);
}

}  // namespace bgls
