// Hash-to-G1 on the device, one message per thread (the step immediately before the
// pairing boundary inside verifyAggSig, bgls/bgls.go:106-111).
//   altbn128 : Keccak-256 (legacy 0x01 padding) try-and-increment, "EVM" variant
//              curves/altbn128.go:509-522, curves/hash.go:53-77
//   bls12-381: blake2b-512(msg || "G1_0"/"G1_1") -> Fouque-Tibouchi / Shallue-van de Woestijne encoding
//              -> cofactor multiplication -> sum of the two points
//              curves/bls12_381.go:349-400, curves/hash.go:79-190
// Output: uncompressed affine big-endian x||y, infinity = zeros.
#pragma once
#include "pairing.cuh"

namespace bgls {

// ---------------------------------------------------------------- Keccak-256 (legacy padding)
HD uint64_t rotl64(uint64_t v, int n) { return (v << n) | (v >> (64 - n)); }

HD void keccak_f1600(uint64_t* a) {
    const uint64_t RC[24] = {0x0000000000000001ull, 0x0000000000008082ull, 0x800000000000808Aull, 0x8000000080008000ull,
                             0x000000000000808Bull, 0x0000000080000001ull, 0x8000000080008081ull, 0x8000000000008009ull,
                             0x000000000000008Aull, 0x0000000000000088ull, 0x0000000080008009ull, 0x000000008000000Aull,
                             0x000000008000808Bull, 0x800000000000008Bull, 0x8000000000008089ull, 0x8000000000008003ull,
                             0x8000000000008002ull, 0x8000000000000080ull, 0x000000000000800Aull, 0x800000008000000Aull,
                             0x8000000080008081ull, 0x8000000000008080ull, 0x0000000080000001ull, 0x8000000080008008ull};
    const int ROT[25] = {0, 1, 62, 28, 27, 36, 44, 6, 55, 20, 3, 10, 43, 25, 39, 41, 45, 15, 21, 8, 18, 2, 61, 56, 14};
    for (int round = 0; round < 24; round++) {
        uint64_t c[5], d[5], b[25];
        for (int x = 0; x < 5; x++) c[x] = a[x] ^ a[x + 5] ^ a[x + 10] ^ a[x + 15] ^ a[x + 20];
        for (int x = 0; x < 5; x++) d[x] = c[(x + 4) % 5] ^ rotl64(c[(x + 1) % 5], 1);
        for (int i = 0; i < 25; i++) a[i] ^= d[i % 5];
        // rho + pi: lane (x, y) -> (y, 2x + 3y); index = x + 5y
        for (int y = 0; y < 5; y++)
            for (int x = 0; x < 5; x++) {
                const int i = x + 5 * y;
                const uint64_t v = ROT[i] ? rotl64(a[i], ROT[i]) : a[i];
                b[y + 5 * ((2 * x + 3 * y) % 5)] = v;
            }
        for (int y = 0; y < 5; y++)
            for (int x = 0; x < 5; x++) a[x + 5 * y] = b[x + 5 * y] ^ ((~b[(x + 1) % 5 + 5 * y]) & b[(x + 2) % 5 + 5 * y]);
        a[0] ^= RC[round];
    }
}
// digest = Keccak-256(prefix || msg)
HD void keccak256_prefixed(uint8_t* digest, uint8_t prefix, const uint8_t* msg, size_t len) {
    uint64_t a[25];
    for (int i = 0; i < 25; i++) a[i] = 0;
    const size_t total = len + 1;
    size_t pos = 0;  // position in the block (bytes)
    for (size_t i = 0; i <= total; i++) {
        uint8_t byte;
        if (i < total) byte = (i == 0) ? prefix : msg[i - 1];
        else byte = 0x01;  // legacy Keccak domain byte
        a[pos >> 3] ^= (uint64_t)byte << (8 * (pos & 7));
        pos++;
        if (i == total) {
            // final bit of the pad10*1 rule goes to the last byte of the block
            a[(136 - 1) >> 3] ^= (uint64_t)0x80 << (8 * ((136 - 1) & 7));
            keccak_f1600(a);
        } else if (pos == 136) {
            keccak_f1600(a);
            pos = 0;
        }
    }
    for (int i = 0; i < 32; i++) digest[i] = (uint8_t)(a[i >> 3] >> (8 * (i & 7)));
}

// ---------------------------------------------------------------- blake2b-512 (unkeyed)
HD uint64_t rotr64(uint64_t v, int n) { return (v >> n) | (v << (64 - n)); }

HD void blake2b_compress(uint64_t* h, const uint8_t* block, uint64_t t, bool last) {
    const uint64_t IV[8] = {0x6a09e667f3bcc908ull, 0xbb67ae8584caa73bull, 0x3c6ef372fe94f82bull, 0xa54ff53a5f1d36f1ull,
                            0x510e527fade682d1ull, 0x9b05688c2b3e6c1full, 0x1f83d9abfb41bd6bull, 0x5be0cd19137e2179ull};
    const uint8_t SIGMA[12][16] = {
        {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3},
        {11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4}, {7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8},
        {9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13}, {2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9},
        {12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11}, {13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10},
        {6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5}, {10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0},
        {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3}};
    uint64_t m[16], v[16];
    for (int i = 0; i < 16; i++) {
        uint64_t w = 0;
        for (int k = 7; k >= 0; k--) w = (w << 8) | block[8 * i + k];
        m[i] = w;
    }
    for (int i = 0; i < 8; i++) { v[i] = h[i]; v[i + 8] = IV[i]; }
    v[12] ^= t;
    if (last) v[14] = ~v[14];
#define BGLS_B2B_G(a, b, c, d, x, y)      \
    v[a] = v[a] + v[b] + (x);             \
    v[d] = rotr64(v[d] ^ v[a], 32);       \
    v[c] = v[c] + v[d];                   \
    v[b] = rotr64(v[b] ^ v[c], 24);       \
    v[a] = v[a] + v[b] + (y);             \
    v[d] = rotr64(v[d] ^ v[a], 16);       \
    v[c] = v[c] + v[d];                   \
    v[b] = rotr64(v[b] ^ v[c], 63);
    for (int r = 0; r < 12; r++) {
        const uint8_t* s = SIGMA[r];
        BGLS_B2B_G(0, 4, 8, 12, m[s[0]], m[s[1]])
        BGLS_B2B_G(1, 5, 9, 13, m[s[2]], m[s[3]])
        BGLS_B2B_G(2, 6, 10, 14, m[s[4]], m[s[5]])
        BGLS_B2B_G(3, 7, 11, 15, m[s[6]], m[s[7]])
        BGLS_B2B_G(0, 5, 10, 15, m[s[8]], m[s[9]])
        BGLS_B2B_G(1, 6, 11, 12, m[s[10]], m[s[11]])
        BGLS_B2B_G(2, 7, 8, 13, m[s[12]], m[s[13]])
        BGLS_B2B_G(3, 4, 9, 14, m[s[14]], m[s[15]])
    }
#undef BGLS_B2B_G
    for (int i = 0; i < 8; i++) h[i] ^= v[i] ^ v[i + 8];
}
// digest(64) = blake2b-512(msg || tag[4])
HD void blake2b512_tagged(uint8_t* digest, const uint8_t* msg, size_t len, const char* tag) {
    uint64_t h[8] = {0x6a09e667f3bcc908ull ^ 0x01010040ull, 0xbb67ae8584caa73bull, 0x3c6ef372fe94f82bull, 0xa54ff53a5f1d36f1ull,
                     0x510e527fade682d1ull, 0x9b05688c2b3e6c1full, 0x1f83d9abfb41bd6bull, 0x5be0cd19137e2179ull};
    const size_t total = len + 4;
    uint8_t block[128];
    size_t done = 0;
    while (true) {
        const size_t remain = total - done;
        const bool last = remain <= 128;
        const size_t take = last ? remain : 128;
        for (size_t i = 0; i < 128; i++) {
            const size_t p = done + i;
            block[i] = i < take ? (p < len ? msg[p] : (uint8_t)tag[p - len]) : 0;
        }
        done += take;
        blake2b_compress(h, block, (uint64_t)done, last);
        if (last) break;
    }
    for (int i = 0; i < 64; i++) digest[i] = (uint8_t)(h[i >> 3] >> (8 * (i & 7)));
}

// ---------------------------------------------------------------- field helpers for the encodings
template <class C> HD void fp_sqrt_candidate(Fp<C>& r, const Fp<C>& a) { fp_pow(r, a, C::PP1D4(), C::N); }  // hash.go:178-190
// isQuadRes (hash.go:254-265: Euler's criterion, 0 counts as a residue) by the binary Jacobi symbol (inv.cuh): the symbol
// of the Montgomery residue a R is that of a because R is a square
template <class C> HD bool fp_is_quad_res(const Fp<C>& a) { return mp_jacobi<C>(a.v) >= 0; }
// parity(x): x > q - x  <=>  canonical x > (q-1)/2                                                           hash.go:169-172
template <class C> HD bool fp_parity(const Fp<C>& a) {
    Fp<C> one, x;
    fp_zero(one);
    one.v[0] = 1;
    fp_mul(x, a, one);  // out of Montgomery form
    const uint32_t* h = C::PM1D2();
    for (int i = C::N - 1; i >= 0; i--) {
        if (x.v[i] != h[i]) return x.v[i] > h[i];
    }
    return false;
}
template <class C> HD void g1_x_to_y2(Fp<C>& r, const Fp<C>& x) {
    Fp<C> t, b;
    fp_sqr(t, x);
    fp_mul(t, t, x);
    fp_set(b, C::B1());
    fp_add(r, t, b);
}
template <class C> HD void g1_store_affine(uint8_t* out, const Fp<C>& x, const Fp<C>& y) {
    fp_to_be<C>(out, x);
    fp_to_be<C>(out + C::FP_BYTES, y);
}

// 32-byte digest (big-endian integer < 2^256, up to ~5.3 q) mod q, in Montgomery form.  The Montgomery product
// assumes operands below q, so the digest is split as T1 * 2^248 + T0 with T0 < 2^248 < q and T1 < 2^8.
template <class C> HD void fp_from_digest32(Fp<C>& r, const uint8_t* d) {
    static_assert(C::FP_BYTES == 32, "written for 32-byte fields");
    uint8_t lo[32], hi[32];
    for (int i = 0; i < 32; i++) { lo[i] = d[i]; hi[i] = 0; }
    lo[0] = 0;
    hi[31] = d[0];
    Fp<C> t1, t0, c;
    fp_from_be<C>(t1, hi);
    fp_from_be<C>(t0, lo);
    fp_set(c, C::TWO248());
    fp_mul(t1, t1, c);
    fp_add(r, t1, t0);
}

// ---------------------------------------------------------------- altbn128: tryAndIncrementEvm (hash.go:53-77)
// one try: keccak(counter || msg) mod q is the abscissa of a curve point?  (px, root) valid when true
template <class C> HD bool bn_hash_try(Fp<C>& px, Fp<C>& root, uint8_t counter, const uint8_t* msg, size_t len) {
    uint8_t h[32];
    keccak256_prefixed(h, counter, msg, len);
    Fp<C> y2, chk;
    fp_from_digest32<C>(px, h);  // h mod q
    g1_x_to_y2(y2, px);
    fp_sqrt_candidate(root, y2);
    fp_sqr(chk, root);
    return fp_eq(chk, y2);
}
// the same test without the square root (Jacobi symbol): which counter the loop stops at
template <class C> HD bool bn_hash_test(uint8_t counter, const uint8_t* msg, size_t len) {
    uint8_t h[32];
    keccak256_prefixed(h, counter, msg, len);
    Fp<C> px, y2;
    fp_from_digest32<C>(px, h);
    g1_x_to_y2(y2, px);
    return fp_is_quad_res(y2);
}
// the sign of y comes from keccak(255 || msg) (hash.go:68-72)
template <class C> HD void bn_hash_finish(uint8_t* out, const Fp<C>& px, Fp<C>& root, const uint8_t* msg, size_t len) {
    uint8_t h[32];
    keccak256_prefixed(h, 255, msg, len);
    if (h[31] & 1) fp_neg(root, root);
    g1_store_affine<C>(out, px, root);
}
template <class C> HDNI void hash_to_g1_keccak_ti(uint8_t* out, const uint8_t* msg, size_t len) {
    uint8_t counter = 0;
    while (true) {
        Fp<C> px, root;
        if (bn_hash_try<C>(px, root, counter, msg, len)) {
            bn_hash_finish<C>(out, px, root, msg, len);
            return;
        }
        counter++;
    }
}

// ---------------------------------------------------------------- bls12-381: Fouque-Tibouchi (hash.go:79-167, bls12_381.go:362-393)
// t: 64 digest bytes (big-endian integer), reduced mod q
template <class C> HD void fp_from_be_wide64(Fp<C>& r, const uint8_t* d) {
    static_assert(C::FP_BYTES == 48, "wide reduction written for 48-byte fields");
    // t = T1 * 2^376 + T0, T0 = low 47 bytes (< 2^376 < q), T1 = high 17 bytes: both operands of the Montgomery
    // products stay below q
    uint8_t hi[48], lo[48];
    for (int i = 0; i < 48; i++) { hi[i] = 0; lo[i] = 0; }
    for (int i = 0; i < 17; i++) hi[31 + i] = d[i];
    for (int i = 0; i < 47; i++) lo[1 + i] = d[17 + i];
    Fp<C> t1, t0, c;
    fp_from_be<C>(t1, hi);
    fp_from_be<C>(t0, lo);
    fp_set(c, C::TWO376());
    fp_mul(t1, t1, c);
    fp_add(r, t1, t0);
}
// sw(): returns an affine point on the curve (never infinity for t != 0, +-sqrt(-5))
template <class C> HD void sw_encode(Fp<C>& x, Fp<C>& y, const Fp<C>& t) {
    Fp<C> one, b, s3, z, A, Bv, AB, inv, w, x0, x1, x2, tmp, y2;
    fp_set(one, C::R1());
    fp_set(b, C::B1());
    fp_set(s3, C::FT_SQRT_NEG3());
    fp_set(z, C::FT_Z());
    // w = sqrt(-3) t / (1 + b + t^2);  1/w^2 is needed for x2: invert A*B once, A = 1+b+t^2, B = sqrt(-3) t
    fp_sqr(A, t);
    fp_add(A, A, one);
    fp_add(A, A, b);
    fp_mul(Bv, s3, t);
    fp_mul(AB, A, Bv);
    fp_inv(inv, AB);
    fp_mul(tmp, Bv, inv);   // 1/A
    fp_mul(w, Bv, tmp);     // B/A
    fp_mul(tmp, t, w);
    fp_sub(x0, z, tmp);     // x0 = z - t w
    fp_neg(x1, x0);
    fp_sub(x1, x1, one);    // x1 = -1 - x0
    // the candidates are tested in the reference's order (hash.go:120-160): x0, then x1, else x2.  The residue tests are
    // Jacobi symbols, so every lane runs exactly ONE square-root exponentiation (three in lockstep before)
    x = x0;
    g1_x_to_y2(y2, x0);
    if (!fp_is_quad_res(y2)) {
        x = x1;
        g1_x_to_y2(y2, x1);
        if (!fp_is_quad_res(y2)) {
            fp_mul(tmp, A, inv);    // 1/B
            fp_mul(tmp, A, tmp);    // A/B = 1/w
            fp_sqr(tmp, tmp);
            fp_add(x2, tmp, one);   // x2 = 1 + 1/w^2
            x = x2;
            g1_x_to_y2(y2, x2);
        }
    }
    fp_sqrt_candidate(y, y2);
    if (fp_parity(y) != fp_parity(t)) fp_neg(y, y);
}
template <class C> HD void ft_point(Jac<Fp<C>>& P, const uint8_t* digest64) {
    Fp<C> t, r1, r2, gx, gy;
    fp_from_be_wide64<C>(t, digest64);
    fp_set(r1, C::FT_ROOT1());
    fp_set(r2, C::FT_ROOT2());
    fp_set(gx, C::G1X());
    fp_set(gy, C::G1Y());
    P.inf = false;
    fe_one(P.Z);
    if (fp_is_zero(t)) { P.inf = true; fp_zero(P.X); fp_zero(P.Y); return; }
    if (fp_eq(t, r1)) { P.X = gx; P.Y = gy; return; }                   // curves/bls12_381.go:386-388
    if (fp_eq(t, r2)) { P.X = gx; fp_neg(P.Y, gy); return; }            // curves/bls12_381.go:388-390
    Jac<Fp<C>> Q;
    Q.inf = false;
    fe_one(Q.Z);
    sw_encode(Q.X, Q.Y, t);
    // cofactor 76329603384216526031706109802092473003 = 0x396c8c005555e1568c00aaab0000aaab
    const uint8_t cof[32] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0x39, 0x6c, 0x8c, 0x00, 0x55, 0x55, 0xe1, 0x56,
                             0x8c, 0x00, 0xaa, 0xab, 0x00, 0x00, 0xaa, 0xab};
    jac_mul(P, Q, cof);
}
// one of the two independent halves of hashToG1BlindingAbstracted (bls12_381.go:362-376): blake2b(msg || tag) -> t ->
// Shallue-van de Woestijne point -> times the cofactor
template <class C> HDNI void ft_half(Jac<Fp<C>>& P, const uint8_t* msg, size_t len, int which) {
    uint8_t d[64];
    blake2b512_tagged(d, msg, len, which ? "G1_1" : "G1_0");
    ft_point<C>(P, d);
}
// Throughput form of the same hash (one thread per message): the cofactor multiplication is a group homomorphism, so
// h Q0 + h Q1 = h (Q0 + Q1) -- ONE 126-bit scalar multiplication per message instead of one per half (it is 70 % of a
// half).  The degenerate inputs that bypass the multiplication in the reference (t = 0 and the two roots,
// bls12_381.go:384-390) take the per-half path below.
template <class C> HD bool ft_special(const Fp<C>& t) {
    Fp<C> r1, r2;
    fp_set(r1, C::FT_ROOT1());
    fp_set(r2, C::FT_ROOT2());
    return fp_is_zero(t) || fp_eq(t, r1) || fp_eq(t, r2);
}
template <class C> HDNI void hash_to_g1_ft(uint8_t* out, const uint8_t* msg, size_t len);
template <class C> HDNI void hash_to_g1_ft_shared_cofactor(uint8_t* out, const uint8_t* msg, size_t len) {
    uint8_t d[64];
    Fp<C> t0, t1;
    blake2b512_tagged(d, msg, len, "G1_0");
    fp_from_be_wide64<C>(t0, d);
    blake2b512_tagged(d, msg, len, "G1_1");
    fp_from_be_wide64<C>(t1, d);
    if (ft_special<C>(t0) || ft_special<C>(t1)) { hash_to_g1_ft<C>(out, msg, len); return; }
    Jac<Fp<C>> Q0, Q1, S, P;
    Q0.inf = Q1.inf = false;
    fe_one(Q0.Z);
    fe_one(Q1.Z);
    sw_encode(Q0.X, Q0.Y, t0);
    sw_encode(Q1.X, Q1.Y, t1);
    jac_add(S, Q0, Q1);
    bool inf;
    jac_to_affine(S.X, S.Y, inf, S);           // affine again: the additions of the scalar multiplication are mixed ones
    if (inf) { for (int i = 0; i < 2 * C::FP_BYTES; i++) out[i] = 0; return; }
    S.inf = false;
    fe_one(S.Z);
    const uint8_t cof[32] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0x39, 0x6c, 0x8c, 0x00, 0x55, 0x55, 0xe1, 0x56,
                             0x8c, 0x00, 0xaa, 0xab, 0x00, 0x00, 0xaa, 0xab};
    jac_mul(P, S, cof);
    jac_store<C>(out, P);
}
template <class C> HDNI void hash_to_g1_ft(uint8_t* out, const uint8_t* msg, size_t len) {
    Jac<Fp<C>> P1, P2, S;
    ft_half<C>(P1, msg, len, 0);
    ft_half<C>(P2, msg, len, 1);
    jac_add(S, P1, P2);
    jac_store<C>(out, S);
}

}  // namespace bgls
