/* TEST INFRASTRUCTURE (oracle) -- not part of the product path.
 *
 * Generic body of the C restatement, included twice by oracle.c:
 *   NL = 4, BN = 1  -> altbn128  (prefix bn_)
 *   NL = 6, BN = 0  -> bls12-381 (prefix bl_)
 * 64-bit limbs, Montgomery form, fully reduced, unsigned __int128 products.
 *
 * The arithmetic below the reference's Pair() (curves/altbn128.go:130-141,
 * curves/bls12_381.go:228-236) lives in un-vendored third-party Go modules
 * (go-ethereum bn256/cloudflare, dis2/bls12); it is restated from the published
 * algorithms: tower Fp2=Fp[i]/(i^2+1), Fp6=Fp2[v]/(v^3-xi), Fp12=Fp6[w]/(w^2-v);
 * optimal-ate Miller loop with homogeneous projective line functions
 * (Costello-Lange-Naehrig), final exponentiation = easy part + exact hard part
 * (BN: Devegili-Scott-Dahab chain; BLS12: ((x-1)^2/3)(x+p)(x^2+p^2-1)+1).
 * Validated against oracle/bgls_oracle.py (tests/test_oracle_c.py).
 */

#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)
#define FN(name) CAT(PFX, name)

typedef uint64_t FN(fp)[NL];
typedef struct { FN(fp) c0, c1; } FN(fp2);          /* c0 + c1*i */
typedef struct { FN(fp2) a0, a1, a2; } FN(fp6);     /* a0 + a1 v + a2 v^2 */
typedef struct { FN(fp6) c0, c1; } FN(fp12);        /* c0 + c1 w */

#define fp FN(fp)
#define fp2 FN(fp2)
#define fp6 FN(fp6)
#define fp12 FN(fp12)

static uint64_t FN(P)[NL];      /* modulus */
static uint64_t FN(N0);         /* -p^-1 mod 2^64 */
static fp FN(R1), FN(R2);       /* R mod p, R^2 mod p */
static fp2 FN(XI);              /* xi in Montgomery form */
static fp2 FN(B2);              /* twist coefficient b' */
static fp2 FN(B2x3);            /* 3 b' */
static fp FN(B1);               /* b */
static fp FN(HALF);             /* 1/2 */
static fp2 FN(GAMMA)[6];        /* gamma_1^k = xi^(k (p-1)/6), k = 0..5 */
static int FN(inited) = 0;

/* ---------------------------------------------------------------- Fp */
static inline int FN(fp_is_zero)(const fp a) {
    uint64_t x = 0;
    for (int i = 0; i < NL; i++) x |= a[i];
    return x == 0;
}
static inline int FN(fp_eq)(const fp a, const fp b) {
    uint64_t x = 0;
    for (int i = 0; i < NL; i++) x |= a[i] ^ b[i];
    return x == 0;
}
static inline void FN(fp_set)(fp r, const fp a) { memcpy(r, a, sizeof(fp)); }
static inline void FN(fp_zero)(fp r) { memset(r, 0, sizeof(fp)); }
static inline int FN(geq_p)(const uint64_t *a) {
    for (int i = NL - 1; i >= 0; i--) {
        if (a[i] > FN(P)[i]) return 1;
        if (a[i] < FN(P)[i]) return 0;
    }
    return 1;
}
static inline void FN(sub_p)(uint64_t *a) {
    uint64_t br = 0;
    for (int i = 0; i < NL; i++) {
        u128 d = (u128)a[i] - FN(P)[i] - br;
        a[i] = (uint64_t)d;
        br = (uint64_t)(d >> 64) & 1;
    }
}
static inline void FN(fp_add)(fp r, const fp a, const fp b) {
    uint64_t c = 0;
    for (int i = 0; i < NL; i++) {
        u128 s = (u128)a[i] + b[i] + c;
        r[i] = (uint64_t)s;
        c = (uint64_t)(s >> 64);
    }
    if (c || FN(geq_p)(r)) FN(sub_p)(r);
}
static inline void FN(fp_sub)(fp r, const fp a, const fp b) {
    uint64_t br = 0;
    for (int i = 0; i < NL; i++) {
        u128 d = (u128)a[i] - b[i] - br;
        r[i] = (uint64_t)d;
        br = (uint64_t)(d >> 64) & 1;
    }
    if (br) {
        uint64_t c = 0;
        for (int i = 0; i < NL; i++) {
            u128 s = (u128)r[i] + FN(P)[i] + c;
            r[i] = (uint64_t)s;
            c = (uint64_t)(s >> 64);
        }
    }
}
static inline void FN(fp_neg)(fp r, const fp a) {
    fp z;
    FN(fp_zero)(z);
    FN(fp_sub)(r, z, a);
}
static inline void FN(fp_dbl)(fp r, const fp a) { FN(fp_add)(r, a, a); }

static void FN(fp_mul)(fp r, const fp a, const fp b) {
    uint64_t t[NL + 2];
    memset(t, 0, sizeof(t));
    orc_fpmul_counter++;
    for (int i = 0; i < NL; i++) {
        uint64_t carry = 0;
        u128 x;
        for (int j = 0; j < NL; j++) {
            x = (u128)a[j] * b[i] + t[j] + carry;
            t[j] = (uint64_t)x;
            carry = (uint64_t)(x >> 64);
        }
        x = (u128)t[NL] + carry;
        t[NL] = (uint64_t)x;
        t[NL + 1] = (uint64_t)(x >> 64);
        uint64_t m = t[0] * FN(N0);
        x = (u128)m * FN(P)[0] + t[0];
        carry = (uint64_t)(x >> 64);
        for (int j = 1; j < NL; j++) {
            x = (u128)m * FN(P)[j] + t[j] + carry;
            t[j - 1] = (uint64_t)x;
            carry = (uint64_t)(x >> 64);
        }
        x = (u128)t[NL] + carry;
        t[NL - 1] = (uint64_t)x;
        t[NL] = t[NL + 1] + (uint64_t)(x >> 64);
    }
    if (t[NL] || FN(geq_p)(t)) FN(sub_p)(t);
    memcpy(r, t, sizeof(fp));
}
static inline void FN(fp_sqr)(fp r, const fp a) { FN(fp_mul)(r, a, a); }

/* r = a^e, e given as nlimbs 64-bit little-endian limbs (plain integer) */
static void FN(fp_pow)(fp r, const fp a, const uint64_t *e, int nlimbs) {
    fp acc, base;
    FN(fp_set)(acc, FN(R1));
    FN(fp_set)(base, a);
    int started = 0;
    for (int i = nlimbs * 64 - 1; i >= 0; i--) {
        if (started) FN(fp_sqr)(acc, acc);
        if ((e[i / 64] >> (i % 64)) & 1) {
            if (started) FN(fp_mul)(acc, acc, base);
            else { FN(fp_set)(acc, base); started = 1; }
        }
    }
    FN(fp_set)(r, acc);
}
static void FN(fp_inv)(fp r, const fp a) {
    uint64_t e[NL];
    memcpy(e, FN(P), sizeof(e));
    e[0] -= 2; /* p is odd and > 2: no borrow */
    FN(fp_pow)(r, a, e, NL);
}
static void FN(fp_from_bytes)(fp r, const uint8_t *be) {
    fp t;
    for (int i = 0; i < NL; i++) {
        uint64_t w = 0;
        for (int k = 0; k < 8; k++) w = (w << 8) | be[(NL - 1 - i) * 8 + k];
        t[i] = w;
    }
    FN(fp_mul)(r, t, FN(R2));
}
static void FN(fp_to_bytes)(uint8_t *be, const fp a) {
    fp one, t;
    FN(fp_zero)(one);
    one[0] = 1;
    FN(fp_mul)(t, a, one);
    for (int i = 0; i < NL; i++)
        for (int k = 0; k < 8; k++) be[(NL - 1 - i) * 8 + k] = (uint8_t)(t[i] >> (56 - 8 * k));
}
static void FN(fp_from_u64)(fp r, uint64_t v) {
    fp t;
    FN(fp_zero)(t);
    t[0] = v;
    FN(fp_mul)(r, t, FN(R2));
}

/* ---------------------------------------------------------------- Fp2 */
static inline void FN(fp2_add)(fp2 *r, const fp2 *a, const fp2 *b) {
    FN(fp_add)(r->c0, a->c0, b->c0);
    FN(fp_add)(r->c1, a->c1, b->c1);
}
static inline void FN(fp2_sub)(fp2 *r, const fp2 *a, const fp2 *b) {
    FN(fp_sub)(r->c0, a->c0, b->c0);
    FN(fp_sub)(r->c1, a->c1, b->c1);
}
static inline void FN(fp2_neg)(fp2 *r, const fp2 *a) {
    FN(fp_neg)(r->c0, a->c0);
    FN(fp_neg)(r->c1, a->c1);
}
static inline void FN(fp2_dbl)(fp2 *r, const fp2 *a) { FN(fp2_add)(r, a, a); }
static inline void FN(fp2_conj)(fp2 *r, const fp2 *a) {
    FN(fp_set)(r->c0, a->c0);
    FN(fp_neg)(r->c1, a->c1);
}
static inline int FN(fp2_is_zero)(const fp2 *a) { return FN(fp_is_zero)(a->c0) && FN(fp_is_zero)(a->c1); }
static inline int FN(fp2_eq)(const fp2 *a, const fp2 *b) { return FN(fp_eq)(a->c0, b->c0) && FN(fp_eq)(a->c1, b->c1); }
static void FN(fp2_mul)(fp2 *r, const fp2 *a, const fp2 *b) {
    fp v0, v1, s, t;
    FN(fp_mul)(v0, a->c0, b->c0);
    FN(fp_mul)(v1, a->c1, b->c1);
    FN(fp_add)(s, a->c0, a->c1);
    FN(fp_add)(t, b->c0, b->c1);
    FN(fp_mul)(s, s, t);
    FN(fp_sub)(s, s, v0);
    FN(fp_sub)(r->c1, s, v1);
    FN(fp_sub)(r->c0, v0, v1);
}
static void FN(fp2_sqr)(fp2 *r, const fp2 *a) {
    fp s, d, m;
    FN(fp_add)(s, a->c0, a->c1);
    FN(fp_sub)(d, a->c0, a->c1);
    FN(fp_mul)(m, a->c0, a->c1);
    FN(fp_mul)(r->c0, s, d);
    FN(fp_dbl)(r->c1, m);
}
static void FN(fp2_mul_fp)(fp2 *r, const fp2 *a, const fp b) {
    FN(fp_mul)(r->c0, a->c0, b);
    FN(fp_mul)(r->c1, a->c1, b);
}
static void FN(fp2_mul_xi)(fp2 *r, const fp2 *a) {
#if BN
    /* (9 + i)(a0 + a1 i) = (9 a0 - a1) + (9 a1 + a0) i */
    fp t0, t1, n0, n1;
    FN(fp_dbl)(t0, a->c0); FN(fp_dbl)(t0, t0); FN(fp_dbl)(t0, t0); FN(fp_add)(t0, t0, a->c0);
    FN(fp_dbl)(t1, a->c1); FN(fp_dbl)(t1, t1); FN(fp_dbl)(t1, t1); FN(fp_add)(t1, t1, a->c1);
    FN(fp_sub)(n0, t0, a->c1);
    FN(fp_add)(n1, t1, a->c0);
    FN(fp_set)(r->c0, n0);
    FN(fp_set)(r->c1, n1);
#else
    /* (1 + i)(a0 + a1 i) = (a0 - a1) + (a0 + a1) i */
    fp n0, n1;
    FN(fp_sub)(n0, a->c0, a->c1);
    FN(fp_add)(n1, a->c0, a->c1);
    FN(fp_set)(r->c0, n0);
    FN(fp_set)(r->c1, n1);
#endif
}
static void FN(fp2_inv)(fp2 *r, const fp2 *a) {
    fp n, t;
    FN(fp_sqr)(n, a->c0);
    FN(fp_sqr)(t, a->c1);
    FN(fp_add)(n, n, t);
    FN(fp_inv)(n, n);
    FN(fp_mul)(r->c0, a->c0, n);
    FN(fp_mul)(t, a->c1, n);
    FN(fp_neg)(r->c1, t);
}
static void FN(fp2_pow)(fp2 *r, const fp2 *a, const uint64_t *e, int nlimbs) {
    fp2 acc, base = *a;
    memset(&acc, 0, sizeof(acc));
    FN(fp_set)(acc.c0, FN(R1));
    for (int i = nlimbs * 64 - 1; i >= 0; i--) {
        FN(fp2_sqr)(&acc, &acc);
        if ((e[i / 64] >> (i % 64)) & 1) FN(fp2_mul)(&acc, &acc, &base);
    }
    *r = acc;
}

/* ---------------------------------------------------------------- Fp6 */
static void FN(fp6_add)(fp6 *r, const fp6 *a, const fp6 *b) {
    FN(fp2_add)(&r->a0, &a->a0, &b->a0);
    FN(fp2_add)(&r->a1, &a->a1, &b->a1);
    FN(fp2_add)(&r->a2, &a->a2, &b->a2);
}
static void FN(fp6_sub)(fp6 *r, const fp6 *a, const fp6 *b) {
    FN(fp2_sub)(&r->a0, &a->a0, &b->a0);
    FN(fp2_sub)(&r->a1, &a->a1, &b->a1);
    FN(fp2_sub)(&r->a2, &a->a2, &b->a2);
}
static void FN(fp6_neg)(fp6 *r, const fp6 *a) {
    FN(fp2_neg)(&r->a0, &a->a0);
    FN(fp2_neg)(&r->a1, &a->a1);
    FN(fp2_neg)(&r->a2, &a->a2);
}
static void FN(fp6_mul_v)(fp6 *r, const fp6 *a) {
    fp2 t;
    FN(fp2_mul_xi)(&t, &a->a2);
    r->a2 = a->a1;
    r->a1 = a->a0;
    r->a0 = t;
}
static void FN(fp6_mul)(fp6 *r, const fp6 *a, const fp6 *b) {
    fp2 v0, v1, v2, s, t, u;
    fp6 o;
    FN(fp2_mul)(&v0, &a->a0, &b->a0);
    FN(fp2_mul)(&v1, &a->a1, &b->a1);
    FN(fp2_mul)(&v2, &a->a2, &b->a2);
    FN(fp2_add)(&s, &a->a1, &a->a2);
    FN(fp2_add)(&t, &b->a1, &b->a2);
    FN(fp2_mul)(&u, &s, &t);
    FN(fp2_sub)(&u, &u, &v1);
    FN(fp2_sub)(&u, &u, &v2);
    FN(fp2_mul_xi)(&u, &u);
    FN(fp2_add)(&o.a0, &v0, &u);
    FN(fp2_add)(&s, &a->a0, &a->a1);
    FN(fp2_add)(&t, &b->a0, &b->a1);
    FN(fp2_mul)(&u, &s, &t);
    FN(fp2_sub)(&u, &u, &v0);
    FN(fp2_sub)(&u, &u, &v1);
    FN(fp2_mul_xi)(&s, &v2);
    FN(fp2_add)(&o.a1, &u, &s);
    FN(fp2_add)(&s, &a->a0, &a->a2);
    FN(fp2_add)(&t, &b->a0, &b->a2);
    FN(fp2_mul)(&u, &s, &t);
    FN(fp2_sub)(&u, &u, &v0);
    FN(fp2_sub)(&u, &u, &v2);
    FN(fp2_add)(&o.a2, &u, &v1);
    *r = o;
}
/* a * (b0 + b1 v) */
static void FN(fp6_mul_by_01)(fp6 *r, const fp6 *a, const fp2 *b0, const fp2 *b1) {
    fp2 v0, v1, s, t, u;
    fp6 o;
    FN(fp2_mul)(&v0, &a->a0, b0);
    FN(fp2_mul)(&v1, &a->a1, b1);
    /* c0 = v0 + xi * (a2*b1) ; (a1+a2)(b1) - v1 = a2 b1 */
    FN(fp2_mul)(&u, &a->a2, b1);
    FN(fp2_mul_xi)(&u, &u);
    FN(fp2_add)(&o.a0, &v0, &u);
    /* c1 = (a0+a1)(b0+b1) - v0 - v1 */
    FN(fp2_add)(&s, &a->a0, &a->a1);
    FN(fp2_add)(&t, b0, b1);
    FN(fp2_mul)(&u, &s, &t);
    FN(fp2_sub)(&u, &u, &v0);
    FN(fp2_sub)(&o.a1, &u, &v1);
    /* c2 = a2 b0 + v1 */
    FN(fp2_mul)(&u, &a->a2, b0);
    FN(fp2_add)(&o.a2, &u, &v1);
    *r = o;
}
/* a * b0 */
static void FN(fp6_mul_by_0)(fp6 *r, const fp6 *a, const fp2 *b0) {
    FN(fp2_mul)(&r->a0, &a->a0, b0);
    FN(fp2_mul)(&r->a1, &a->a1, b0);
    FN(fp2_mul)(&r->a2, &a->a2, b0);
}
/* a * (b1 v) */
static void FN(fp6_mul_by_1)(fp6 *r, const fp6 *a, const fp2 *b1) {
    fp2 t0, t1, t2;
    FN(fp2_mul)(&t0, &a->a2, b1);
    FN(fp2_mul_xi)(&t0, &t0);
    FN(fp2_mul)(&t1, &a->a0, b1);
    FN(fp2_mul)(&t2, &a->a1, b1);
    r->a0 = t0;
    r->a1 = t1;
    r->a2 = t2;
}
static void FN(fp6_inv)(fp6 *r, const fp6 *a) {
    fp2 t0, t1, t2, s, d;
    FN(fp2_sqr)(&t0, &a->a0);
    FN(fp2_mul)(&s, &a->a1, &a->a2);
    FN(fp2_mul_xi)(&s, &s);
    FN(fp2_sub)(&t0, &t0, &s);
    FN(fp2_sqr)(&t1, &a->a2);
    FN(fp2_mul_xi)(&t1, &t1);
    FN(fp2_mul)(&s, &a->a0, &a->a1);
    FN(fp2_sub)(&t1, &t1, &s);
    FN(fp2_sqr)(&t2, &a->a1);
    FN(fp2_mul)(&s, &a->a0, &a->a2);
    FN(fp2_sub)(&t2, &t2, &s);
    FN(fp2_mul)(&d, &a->a2, &t1);
    FN(fp2_mul)(&s, &a->a1, &t2);
    FN(fp2_add)(&d, &d, &s);
    FN(fp2_mul_xi)(&d, &d);
    FN(fp2_mul)(&s, &a->a0, &t0);
    FN(fp2_add)(&d, &d, &s);
    FN(fp2_inv)(&d, &d);
    FN(fp2_mul)(&r->a0, &t0, &d);
    FN(fp2_mul)(&r->a1, &t1, &d);
    FN(fp2_mul)(&r->a2, &t2, &d);
}

/* ---------------------------------------------------------------- Fp12 */
static void FN(fp12_one)(fp12 *r) {
    memset(r, 0, sizeof(*r));
    FN(fp_set)(r->c0.a0.c0, FN(R1));
}
static int FN(fp12_eq)(const fp12 *a, const fp12 *b) { return memcmp(a, b, sizeof(*a)) == 0; }
static void FN(fp12_mul)(fp12 *r, const fp12 *a, const fp12 *b) {
    fp6 t0, t1, s, t, u;
    FN(fp6_mul)(&t0, &a->c0, &b->c0);
    FN(fp6_mul)(&t1, &a->c1, &b->c1);
    FN(fp6_add)(&s, &a->c0, &a->c1);
    FN(fp6_add)(&t, &b->c0, &b->c1);
    FN(fp6_mul)(&u, &s, &t);
    FN(fp6_sub)(&u, &u, &t0);
    FN(fp6_sub)(&r->c1, &u, &t1);
    FN(fp6_mul_v)(&t1, &t1);
    FN(fp6_add)(&r->c0, &t0, &t1);
}
static void FN(fp12_sqr)(fp12 *r, const fp12 *a) {
    fp6 t, s, u, vt;
    FN(fp6_mul)(&t, &a->c0, &a->c1);
    FN(fp6_add)(&s, &a->c0, &a->c1);
    FN(fp6_mul_v)(&u, &a->c1);
    FN(fp6_add)(&u, &u, &a->c0);
    FN(fp6_mul)(&s, &s, &u);
    FN(fp6_mul_v)(&vt, &t);
    FN(fp6_sub)(&s, &s, &t);
    FN(fp6_sub)(&r->c0, &s, &vt);
    FN(fp6_add)(&r->c1, &t, &t);
}
static void FN(fp12_conj)(fp12 *r, const fp12 *a) {
    r->c0 = a->c0;
    FN(fp6_neg)(&r->c1, &a->c1);
}
static void FN(fp12_inv)(fp12 *r, const fp12 *a) {
    fp6 t0, t1;
    FN(fp6_mul)(&t0, &a->c0, &a->c0);
    FN(fp6_mul)(&t1, &a->c1, &a->c1);
    FN(fp6_mul_v)(&t1, &t1);
    FN(fp6_sub)(&t0, &t0, &t1);
    FN(fp6_inv)(&t0, &t0);
    FN(fp6_mul)(&r->c0, &a->c0, &t0);
    FN(fp6_mul)(&t1, &a->c1, &t0);
    FN(fp6_neg)(&r->c1, &t1);
}
/* coefficient of w^k: even k -> c0.a[k/2], odd k -> c1.a[(k-1)/2] */
static fp2 *FN(fp12_coef)(fp12 *a, int k) {
    fp6 *h = (k & 1) ? &a->c1 : &a->c0;
    int j = k >> 1;
    return j == 0 ? &h->a0 : (j == 1 ? &h->a1 : &h->a2);
}
/* a^p: coefficient k -> conj(c_k) * gamma_1^k  (w^(p-1) = xi^((p-1)/6)) */
static void FN(fp12_frob)(fp12 *r, const fp12 *a) {
    fp12 t = *a;
    for (int k = 0; k < 6; k++) {
        fp2 *c = FN(fp12_coef)(&t, k);
        fp2 cc;
        FN(fp2_conj)(&cc, c);
        if (k == 0) *c = cc;
        else FN(fp2_mul)(c, &cc, &FN(GAMMA)[k]);
    }
    *r = t;
}
/* Granger-Scott squaring, valid in the cyclotomic subgroup only */
static void FN(fp4_sqr)(fp2 *o0, fp2 *o1, const fp2 *a, const fp2 *b) {
    fp2 t0, t1, s;
    FN(fp2_sqr)(&t0, a);
    FN(fp2_sqr)(&t1, b);
    FN(fp2_add)(&s, a, b);
    FN(fp2_sqr)(&s, &s);
    FN(fp2_sub)(&s, &s, &t0);
    FN(fp2_sub)(o1, &s, &t1);
    FN(fp2_mul_xi)(&t1, &t1);
    FN(fp2_add)(o0, &t1, &t0);
}
static void FN(fp12_cyc_sqr)(fp12 *r, const fp12 *f) {
    fp2 z0 = f->c0.a0, z4 = f->c0.a1, z3 = f->c0.a2, z2 = f->c1.a0, z1 = f->c1.a1, z5 = f->c1.a2;
    fp2 t0, t1, t2, t3, s;
    FN(fp4_sqr)(&t0, &t1, &z0, &z1);
    FN(fp2_sub)(&s, &t0, &z0); FN(fp2_dbl)(&s, &s); FN(fp2_add)(&z0, &s, &t0);
    FN(fp2_add)(&s, &t1, &z1); FN(fp2_dbl)(&s, &s); FN(fp2_add)(&z1, &s, &t1);
    FN(fp4_sqr)(&t0, &t1, &z2, &z3);
    FN(fp4_sqr)(&t2, &t3, &z4, &z5);
    FN(fp2_sub)(&s, &t0, &z4); FN(fp2_dbl)(&s, &s); FN(fp2_add)(&z4, &s, &t0);
    FN(fp2_add)(&s, &t1, &z5); FN(fp2_dbl)(&s, &s); FN(fp2_add)(&z5, &s, &t1);
    FN(fp2_mul_xi)(&t0, &t3);
    FN(fp2_add)(&s, &t0, &z2); FN(fp2_dbl)(&s, &s); FN(fp2_add)(&z2, &s, &t0);
    FN(fp2_sub)(&s, &t2, &z3); FN(fp2_dbl)(&s, &s); FN(fp2_add)(&z3, &s, &t2);
    r->c0.a0 = z0; r->c0.a1 = z4; r->c0.a2 = z3;
    r->c1.a0 = z2; r->c1.a1 = z1; r->c1.a2 = z5;
}
/* a^e for unitary a, e a 64-bit-limb integer (plain square and multiply, cyclotomic squarings) */
static void FN(fp12_cyc_pow)(fp12 *r, const fp12 *a, const uint64_t *e, int nlimbs) {
    fp12 acc, base = *a;
    int started = 0;
    FN(fp12_one)(&acc);
    for (int i = nlimbs * 64 - 1; i >= 0; i--) {
        if (started) FN(fp12_cyc_sqr)(&acc, &acc);
        if ((e[i / 64] >> (i % 64)) & 1) {
            if (started) FN(fp12_mul)(&acc, &acc, &base);
            else { acc = base; started = 1; }
        }
    }
    *r = acc;
}
static void FN(fp12_pow)(fp12 *r, const fp12 *a, const uint64_t *e, int nlimbs) {
    fp12 acc, base = *a;
    FN(fp12_one)(&acc);
    for (int i = nlimbs * 64 - 1; i >= 0; i--) {
        FN(fp12_sqr)(&acc, &acc);
        if ((e[i / 64] >> (i % 64)) & 1) FN(fp12_mul)(&acc, &acc, &base);
    }
    *r = acc;
}
/* GT byte layout: w-powers 5,3,1,4,2,0, each (im, re), big-endian (see bgls_oracle.marshal_gt) */
static const int FN(GT_ORDER)[6] = {5, 3, 1, 4, 2, 0};
static void FN(fp12_to_bytes)(uint8_t *out, const fp12 *a) {
    fp12 t = *a;
    for (int i = 0; i < 6; i++) {
        fp2 *c = FN(fp12_coef)(&t, FN(GT_ORDER)[i]);
        FN(fp_to_bytes)(out + (2 * i) * NL * 8, c->c1);
        FN(fp_to_bytes)(out + (2 * i + 1) * NL * 8, c->c0);
    }
}
static void FN(fp12_from_bytes)(fp12 *a, const uint8_t *in) {
    for (int i = 0; i < 6; i++) {
        fp2 *c = FN(fp12_coef)(a, FN(GT_ORDER)[i]);
        FN(fp_from_bytes)(c->c1, in + (2 * i) * NL * 8);
        FN(fp_from_bytes)(c->c0, in + (2 * i + 1) * NL * 8);
    }
}

/* ---------------------------------------------------------------- sparse line multiplication */
#if BN
/* D-type: line = l0 + l1 w + l3 w^3  ->  d0 = (l0,0,0), d1 = (l1,l3,0) */
static void FN(fp12_mul_line)(fp12 *f, const fp2 *l0, const fp2 *l1, const fp2 *l3) {
    fp6 t0, t1, s, u;
    fp2 e0;
    FN(fp6_mul_by_0)(&t0, &f->c0, l0);
    FN(fp6_mul_by_01)(&t1, &f->c1, l1, l3);
    FN(fp6_add)(&s, &f->c0, &f->c1);
    FN(fp2_add)(&e0, l0, l1);
    FN(fp6_mul_by_01)(&u, &s, &e0, l3);
    FN(fp6_sub)(&u, &u, &t0);
    FN(fp6_sub)(&f->c1, &u, &t1);
    FN(fp6_mul_v)(&t1, &t1);
    FN(fp6_add)(&f->c0, &t0, &t1);
}
#else
/* M-type: line = l0 + l2 w^2 + l3 w^3 ->  d0 = (l0,l2,0), d1 = (0,l3,0) */
static void FN(fp12_mul_line)(fp12 *f, const fp2 *l0, const fp2 *l2, const fp2 *l3) {
    fp6 t0, t1, s, u;
    fp2 e1;
    FN(fp6_mul_by_01)(&t0, &f->c0, l0, l2);
    FN(fp6_mul_by_1)(&t1, &f->c1, l3);
    FN(fp6_add)(&s, &f->c0, &f->c1);
    FN(fp2_add)(&e1, l2, l3);
    FN(fp6_mul_by_01)(&u, &s, l0, &e1);
    FN(fp6_sub)(&u, &u, &t0);
    FN(fp6_sub)(&f->c1, &u, &t1);
    FN(fp6_mul_v)(&t1, &t1);
    FN(fp6_add)(&f->c0, &t0, &t1);
}
#endif

/* ---------------------------------------------------------------- Miller loop */
typedef struct { fp2 X, Y, Z; } FN(g2proj);
typedef struct { fp x, y; int inf; } FN(g1aff);
typedef struct { fp2 x, y; int inf; } FN(g2aff);
#define g2proj FN(g2proj)
#define g1aff FN(g1aff)
#define g2aff FN(g2aff)

/* T <- 2T, f <- f * l_{T,T}(P).  Homogeneous projective, a = 0.
 *   H = 2YZ, B = Y^2, E = 3b'Z^2:  line = (H yP) + (-3X^2 xP) [w or w^2] + (B - E) [w^3 or 1] */
static void FN(dbl_step)(fp12 *f, g2proj *T, const g1aff *P) {
    fp2 A, B, C, E, F, G, H, X2, t, l_y, l_x, l_c;
    const uint64_t *half = FN(HALF);
    FN(fp2_mul)(&A, &T->X, &T->Y);
    FN(fp2_mul_fp)(&A, &A, half);
    FN(fp2_sqr)(&B, &T->Y);
    FN(fp2_sqr)(&C, &T->Z);
    FN(fp2_mul)(&E, &C, &FN(B2x3));
    FN(fp2_dbl)(&F, &E);
    FN(fp2_add)(&F, &F, &E);
    FN(fp2_add)(&G, &B, &F);
    FN(fp2_mul_fp)(&G, &G, half);
    FN(fp2_add)(&H, &T->Y, &T->Z);
    FN(fp2_sqr)(&H, &H);
    FN(fp2_sub)(&H, &H, &B);
    FN(fp2_sub)(&H, &H, &C);
    FN(fp2_sqr)(&X2, &T->X);
    /* line */
    FN(fp2_mul_fp)(&l_y, &H, P->y);
    FN(fp2_dbl)(&t, &X2);
    FN(fp2_add)(&t, &t, &X2);
    FN(fp2_mul_fp)(&l_x, &t, P->x);
    FN(fp2_neg)(&l_x, &l_x);
    FN(fp2_sub)(&l_c, &B, &E);
    /* point */
    FN(fp2_sub)(&t, &B, &F);
    FN(fp2_mul)(&T->X, &A, &t);
    FN(fp2_sqr)(&t, &G);
    FN(fp2_sqr)(&C, &E);
    FN(fp2_dbl)(&A, &C);
    FN(fp2_add)(&A, &A, &C);
    FN(fp2_sub)(&T->Y, &t, &A);
    FN(fp2_mul)(&T->Z, &B, &H);
#if BN
    FN(fp12_mul_line)(f, &l_y, &l_x, &l_c);
#else
    FN(fp12_mul_line)(f, &l_c, &l_x, &l_y);
#endif
}
/* T <- T + Q (Q affine), f <- f * l_{T,Q}(P).
 *   theta = Y - yQ Z, lam = X - xQ Z: line = (lam yP) + (-theta xP)[..] + (theta xQ - lam yQ)[..] */
static void FN(add_step)(fp12 *f, g2proj *T, const g2aff *Q, const g1aff *P) {
    fp2 th, la, C, D, E, F, G, H, t, l_y, l_x, l_c;
    FN(fp2_mul)(&t, &Q->y, &T->Z);
    FN(fp2_sub)(&th, &T->Y, &t);
    FN(fp2_mul)(&t, &Q->x, &T->Z);
    FN(fp2_sub)(&la, &T->X, &t);
    FN(fp2_mul_fp)(&l_y, &la, P->y);
    FN(fp2_mul_fp)(&l_x, &th, P->x);
    FN(fp2_neg)(&l_x, &l_x);
    FN(fp2_mul)(&l_c, &th, &Q->x);
    FN(fp2_mul)(&t, &la, &Q->y);
    FN(fp2_sub)(&l_c, &l_c, &t);
    FN(fp2_sqr)(&C, &th);
    FN(fp2_sqr)(&D, &la);
    FN(fp2_mul)(&E, &la, &D);
    FN(fp2_mul)(&F, &T->Z, &C);
    FN(fp2_mul)(&G, &T->X, &D);
    FN(fp2_add)(&H, &E, &F);
    FN(fp2_sub)(&H, &H, &G);
    FN(fp2_sub)(&H, &H, &G);
    FN(fp2_mul)(&T->X, &la, &H);
    FN(fp2_sub)(&t, &G, &H);
    FN(fp2_mul)(&t, &th, &t);
    FN(fp2_mul)(&G, &E, &T->Y);
    FN(fp2_sub)(&T->Y, &t, &G);
    FN(fp2_mul)(&T->Z, &T->Z, &E);
#if BN
    FN(fp12_mul_line)(f, &l_y, &l_x, &l_c);
#else
    FN(fp12_mul_line)(f, &l_c, &l_x, &l_y);
#endif
}

static void FN(miller)(fp12 *f, const g1aff *P, const g2aff *Q) {
    FN(fp12_one)(f);
    if (P->inf || Q->inf) return;
    g2proj T;
    T.X = Q->x;
    T.Y = Q->y;
    memset(&T.Z, 0, sizeof(T.Z));
    FN(fp_set)(T.Z.c0, FN(R1));
#if BN
    /* 6u+2, u = 4965661367192848881: 65 bits */
    const uint64_t s_lo = 0x9D797039BE763BA8ull; /* low 64 bits of 29793968203157093288 */
    const int top = 64;                           /* bit 64 is the leading one */
    for (int i = top - 1; i >= 0; i--) {
        FN(fp12_sqr)(f, f);
        FN(dbl_step)(f, &T, P);
        if ((s_lo >> i) & 1) FN(add_step)(f, &T, Q, P);
    }
    g2aff Q1, Q2;
    Q1.inf = Q2.inf = 0;
    fp2 t;
    FN(fp2_conj)(&t, &Q->x); FN(fp2_mul)(&Q1.x, &t, &FN(GAMMA)[2]);
    FN(fp2_conj)(&t, &Q->y); FN(fp2_mul)(&Q1.y, &t, &FN(GAMMA)[3]);
    FN(fp2_conj)(&t, &Q1.x); FN(fp2_mul)(&Q2.x, &t, &FN(GAMMA)[2]);
    FN(fp2_conj)(&t, &Q1.y); FN(fp2_mul)(&Q2.y, &t, &FN(GAMMA)[3]);
    FN(fp2_neg)(&Q2.y, &Q2.y);
    FN(add_step)(f, &T, &Q1, P);
    FN(add_step)(f, &T, &Q2, P);
#else
    const uint64_t x = 0xD201000000010000ull;
    for (int i = 62; i >= 0; i--) {
        FN(fp12_sqr)(f, f);
        FN(dbl_step)(f, &T, P);
        if ((x >> i) & 1) FN(add_step)(f, &T, Q, P);
    }
    FN(fp12_conj)(f, f);
#endif
}

/* ---------------------------------------------------------------- final exponentiation */
static void FN(final_exp_easy)(fp12 *r, const fp12 *f) {
    fp12 t0, t1;
    FN(fp12_conj)(&t0, f);
    FN(fp12_inv)(&t1, f);
    FN(fp12_mul)(&t0, &t0, &t1);          /* f^(p^6-1) */
    FN(fp12_frob)(&t1, &t0);
    FN(fp12_frob)(&t1, &t1);
    FN(fp12_mul)(r, &t1, &t0);            /* ^(p^2+1) */
}
#if BN
static void FN(exp_u)(fp12 *r, const fp12 *a) {
    const uint64_t u = 4965661367192848881ull;
    FN(fp12_cyc_pow)(r, a, &u, 1);
}
static void FN(final_exp)(fp12 *r, const fp12 *f) {
    fp12 t1, t0, fp_, fp2_, fp3_, fu, fu2, fu3, y0, y1, y2, y3, y4, y5, y6, fu2p, fu3p;
    FN(final_exp_easy)(&t1, f);
    FN(fp12_frob)(&fp_, &t1);
    FN(fp12_frob)(&fp2_, &fp_);
    FN(fp12_frob)(&fp3_, &fp2_);
    FN(exp_u)(&fu, &t1);
    FN(exp_u)(&fu2, &fu);
    FN(exp_u)(&fu3, &fu2);
    FN(fp12_frob)(&y3, &fu);
    FN(fp12_frob)(&fu2p, &fu2);
    FN(fp12_frob)(&fu3p, &fu3);
    FN(fp12_frob)(&y2, &fu2p);
    FN(fp12_mul)(&y0, &fp_, &fp2_);
    FN(fp12_mul)(&y0, &y0, &fp3_);
    FN(fp12_conj)(&y1, &t1);
    FN(fp12_conj)(&y5, &fu2);
    FN(fp12_conj)(&y3, &y3);
    FN(fp12_mul)(&y4, &fu, &fu2p);
    FN(fp12_conj)(&y4, &y4);
    FN(fp12_mul)(&y6, &fu3, &fu3p);
    FN(fp12_conj)(&y6, &y6);
    FN(fp12_cyc_sqr)(&t0, &y6);
    FN(fp12_mul)(&t0, &t0, &y4);
    FN(fp12_mul)(&t0, &t0, &y5);
    FN(fp12_mul)(&t1, &y3, &y5);
    FN(fp12_mul)(&t1, &t1, &t0);
    FN(fp12_mul)(&t0, &t0, &y2);
    FN(fp12_cyc_sqr)(&t1, &t1);
    FN(fp12_mul)(&t1, &t1, &t0);
    FN(fp12_cyc_sqr)(&t1, &t1);
    FN(fp12_mul)(&t0, &t1, &y1);
    FN(fp12_mul)(&t1, &t1, &y0);
    FN(fp12_cyc_sqr)(&t0, &t0);
    FN(fp12_mul)(r, &t0, &t1);
}
#else
static void FN(exp_absx)(fp12 *r, const fp12 *a) {
    const uint64_t x = 0xD201000000010000ull;
    FN(fp12_cyc_pow)(r, a, &x, 1);
}
static void FN(final_exp)(fp12 *r, const fp12 *f) {
    /* hard = ((x-1)^2/3) (x+p) (x^2+p^2-1) + 1, x = -|x| */
    const uint64_t c[2] = {0x8c00aaab0000aaabull, 0x396c8c005555e156ull}; /* (|x|+1)^2/3 */
    fp12 m, y0, y1, y2, t, s;
    FN(final_exp_easy)(&m, f);
    FN(fp12_cyc_pow)(&y0, &m, c, 2);
    FN(exp_absx)(&t, &y0);
    FN(fp12_conj)(&t, &t);                /* y0^x */
    FN(fp12_frob)(&s, &y0);
    FN(fp12_mul)(&y1, &t, &s);            /* y0^(x+p) */
    FN(exp_absx)(&t, &y1);
    FN(exp_absx)(&t, &t);                 /* y1^(x^2) */
    FN(fp12_frob)(&s, &y1);
    FN(fp12_frob)(&s, &s);
    FN(fp12_mul)(&y2, &t, &s);
    FN(fp12_conj)(&t, &y1);
    FN(fp12_mul)(&y2, &y2, &t);           /* y1^(x^2+p^2-1) */
    FN(fp12_mul)(r, &y2, &m);
}
#endif
/* definition check: f^((p^12-1)/r) with the exponent supplied by the caller (python computes it) */
static void FN(final_exp_slow)(fp12 *r, const fp12 *f, const uint64_t *e, int nlimbs) {
    FN(fp12_pow)(r, f, e, nlimbs);
}

/* ---------------------------------------------------------------- generic Jacobian group law (k = 1: G1, k = 2: G2) */
typedef struct { uint64_t v[3][2 * NL]; int inf; } FN(jac); /* X,Y,Z; each k*NL limbs */
#define jac FN(jac)
#define FE uint64_t *
static void FN(fe_add)(int k, FE r, const uint64_t *a, const uint64_t *b) { for (int i = 0; i < k; i++) FN(fp_add)(r + i * NL, a + i * NL, b + i * NL); }
static void FN(fe_sub)(int k, FE r, const uint64_t *a, const uint64_t *b) { for (int i = 0; i < k; i++) FN(fp_sub)(r + i * NL, a + i * NL, b + i * NL); }
static void FN(fe_mul)(int k, FE r, const uint64_t *a, const uint64_t *b) {
    if (k == 1) FN(fp_mul)(r, a, b);
    else { fp2 o; FN(fp2_mul)(&o, (const fp2 *)a, (const fp2 *)b); memcpy(r, &o, sizeof(o)); }
}
static void FN(fe_inv)(int k, FE r, const uint64_t *a) {
    if (k == 1) FN(fp_inv)(r, a);
    else { fp2 o; FN(fp2_inv)(&o, (const fp2 *)a); memcpy(r, &o, sizeof(o)); }
}
static int FN(fe_is_zero)(int k, const uint64_t *a) { for (int i = 0; i < k * NL; i++) if (a[i]) return 0; return 1; }
static int FN(fe_eq)(int k, const uint64_t *a, const uint64_t *b) { return memcmp(a, b, k * NL * 8) == 0; }

static void FN(jac_dbl)(int k, jac *r, const jac *p) {
    if (p->inf || FN(fe_is_zero)(k, p->v[1])) { r->inf = 1; return; }
    uint64_t A[2 * NL], B[2 * NL], C[2 * NL], D[2 * NL], E[2 * NL], F[2 * NL], t[2 * NL];
    jac o; o.inf = 0;
    FN(fe_mul)(k, A, p->v[0], p->v[0]);
    FN(fe_mul)(k, B, p->v[1], p->v[1]);
    FN(fe_mul)(k, C, B, B);
    FN(fe_add)(k, t, p->v[0], B);
    FN(fe_mul)(k, t, t, t);
    FN(fe_sub)(k, t, t, A);
    FN(fe_sub)(k, t, t, C);
    FN(fe_add)(k, D, t, t);
    FN(fe_add)(k, E, A, A);
    FN(fe_add)(k, E, E, A);
    FN(fe_mul)(k, F, E, E);
    FN(fe_sub)(k, t, F, D);
    FN(fe_sub)(k, o.v[0], t, D);
    FN(fe_mul)(k, t, p->v[1], p->v[2]);
    FN(fe_add)(k, o.v[2], t, t);
    FN(fe_sub)(k, t, D, o.v[0]);
    FN(fe_mul)(k, t, E, t);
    FN(fe_add)(k, C, C, C);
    FN(fe_add)(k, C, C, C);
    FN(fe_add)(k, C, C, C);
    FN(fe_sub)(k, o.v[1], t, C);
    *r = o;
}
static void FN(jac_add)(int k, jac *r, const jac *p, const jac *q) {
    if (p->inf) { *r = *q; return; }
    if (q->inf) { *r = *p; return; }
    uint64_t Z1Z1[2 * NL], Z2Z2[2 * NL], U1[2 * NL], U2[2 * NL], S1[2 * NL], S2[2 * NL], H[2 * NL], R[2 * NL], t[2 * NL], HH[2 * NL], HHH[2 * NL], V[2 * NL];
    FN(fe_mul)(k, Z1Z1, p->v[2], p->v[2]);
    FN(fe_mul)(k, Z2Z2, q->v[2], q->v[2]);
    FN(fe_mul)(k, U1, p->v[0], Z2Z2);
    FN(fe_mul)(k, U2, q->v[0], Z1Z1);
    FN(fe_mul)(k, t, q->v[2], Z2Z2);
    FN(fe_mul)(k, S1, p->v[1], t);
    FN(fe_mul)(k, t, p->v[2], Z1Z1);
    FN(fe_mul)(k, S2, q->v[1], t);
    if (FN(fe_eq)(k, U1, U2)) {
        if (FN(fe_eq)(k, S1, S2)) { FN(jac_dbl)(k, r, p); return; }
        r->inf = 1;
        return;
    }
    jac o; o.inf = 0;
    FN(fe_sub)(k, H, U2, U1);
    FN(fe_sub)(k, R, S2, S1);
    FN(fe_mul)(k, HH, H, H);
    FN(fe_mul)(k, HHH, HH, H);
    FN(fe_mul)(k, V, U1, HH);
    FN(fe_mul)(k, t, R, R);
    FN(fe_sub)(k, t, t, HHH);
    FN(fe_sub)(k, t, t, V);
    FN(fe_sub)(k, o.v[0], t, V);
    FN(fe_sub)(k, t, V, o.v[0]);
    FN(fe_mul)(k, t, R, t);
    FN(fe_mul)(k, S1, S1, HHH);
    FN(fe_sub)(k, o.v[1], t, S1);
    FN(fe_mul)(k, t, p->v[2], q->v[2]);
    FN(fe_mul)(k, o.v[2], t, H);
    *r = o;
}
static void FN(jac_from_bytes)(int k, jac *r, const uint8_t *in) {
    const int fb = NL * 8;
    int allzero = 1;
    for (int i = 0; i < 2 * k * fb; i++) if (in[i]) { allzero = 0; break; }
    memset(r, 0, sizeof(*r));
    if (allzero || (!BN && (in[0] & 0x40))) { r->inf = 1; return; }
    if (k == 1) {
        FN(fp_from_bytes)(r->v[0], in);
        FN(fp_from_bytes)(r->v[1], in + fb);
    } else { /* x_im, x_re, y_im, y_re */
        FN(fp_from_bytes)(r->v[0] + NL, in);
        FN(fp_from_bytes)(r->v[0], in + fb);
        FN(fp_from_bytes)(r->v[1] + NL, in + 2 * fb);
        FN(fp_from_bytes)(r->v[1], in + 3 * fb);
    }
    FN(fp_set)(r->v[2], FN(R1));
}
static void FN(jac_to_bytes)(int k, uint8_t *out, const jac *p) {
    const int fb = NL * 8;
    memset(out, 0, 2 * k * fb);
    if (p->inf || FN(fe_is_zero)(k, p->v[2])) return;
    uint64_t zi[2 * NL], zi2[2 * NL], zi3[2 * NL], x[2 * NL], y[2 * NL];
    FN(fe_inv)(k, zi, p->v[2]);
    FN(fe_mul)(k, zi2, zi, zi);
    FN(fe_mul)(k, zi3, zi2, zi);
    FN(fe_mul)(k, x, p->v[0], zi2);
    FN(fe_mul)(k, y, p->v[1], zi3);
    if (k == 1) {
        FN(fp_to_bytes)(out, x);
        FN(fp_to_bytes)(out + fb, y);
    } else {
        FN(fp_to_bytes)(out, x + NL);
        FN(fp_to_bytes)(out + fb, x);
        FN(fp_to_bytes)(out + 2 * fb, y + NL);
        FN(fp_to_bytes)(out + 3 * fb, y);
    }
}
/* scalar: 32 bytes big-endian, non-negative */
static void FN(jac_mul)(int k, jac *r, const jac *p, const uint8_t *scalar32) {
    jac acc; memset(&acc, 0, sizeof(acc)); acc.inf = 1;
    for (int i = 0; i < 256; i++) {
        FN(jac_dbl)(k, &acc, &acc);
        if ((scalar32[i / 8] >> (7 - (i % 8))) & 1) FN(jac_add)(k, &acc, &acc, p);
    }
    *r = acc;
}
static int FN(on_curve)(int k, const jac *p) {
    if (p->inf) return 1;
    uint64_t l[2 * NL], rr[2 * NL], t[2 * NL];
    memset(t, 0, sizeof(t));
    FN(fe_mul)(k, l, p->v[1], p->v[1]);
    FN(fe_mul)(k, rr, p->v[0], p->v[0]);
    FN(fe_mul)(k, rr, rr, p->v[0]);
    if (k == 1) FN(fp_add)(rr, rr, FN(B1));
    else FN(fe_add)(k, rr, rr, (const uint64_t *)&FN(B2));
    return FN(fe_eq)(k, l, rr);
}

/* ---------------------------------------------------------------- init */
static void FN(limbs_from_hex)(uint64_t *out, int n, const char *hex) {
    memset(out, 0, n * 8);
    int len = (int)strlen(hex);
    for (int i = 0; i < len; i++) {
        char ch = hex[len - 1 - i];
        uint64_t d = (ch >= '0' && ch <= '9') ? ch - '0' : (ch | 32) - 'a' + 10;
        out[i / 16] |= d << (4 * (i % 16));
    }
}
/* q = a / d for small d, returns remainder */
static uint64_t FN(limbs_div_small)(uint64_t *q, const uint64_t *a, int n, uint64_t d) {
    u128 rem = 0;
    for (int i = n - 1; i >= 0; i--) {
        u128 cur = (rem << 64) | a[i];
        q[i] = (uint64_t)(cur / d);
        rem = cur % d;
    }
    return (uint64_t)rem;
}
static void FN(init)(void) {
    if (FN(inited)) return;
#if BN
    FN(limbs_from_hex)(FN(P), NL, "30644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd47");
#else
    FN(limbs_from_hex)(FN(P), NL, "1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab");
#endif
    uint64_t inv = 1;
    for (int i = 0; i < 6; i++) inv *= 2 - FN(P)[0] * inv;
    FN(N0) = (uint64_t)0 - inv;
    /* R mod p by 64*NL modular doublings of 1; R^2 by 64*NL more */
    fp x;
    FN(fp_zero)(x);
    x[0] = 1;
    for (int i = 0; i < 64 * NL; i++) FN(fp_add)(x, x, x);
    FN(fp_set)(FN(R1), x);
    for (int i = 0; i < 64 * NL; i++) FN(fp_add)(x, x, x);
    FN(fp_set)(FN(R2), x);
    memset(&FN(XI), 0, sizeof(fp2));
    memset(&FN(B2), 0, sizeof(fp2));
#if BN
    FN(fp_from_u64)(FN(XI).c0, 9);
    FN(fp_from_u64)(FN(XI).c1, 1);
    FN(fp_from_u64)(FN(B1), 3);
    { /* b' = 3 / xi */
        fp2 t, b;
        memset(&b, 0, sizeof(b));
        FN(fp_from_u64)(b.c0, 3);
        FN(fp2_inv)(&t, &FN(XI));
        FN(fp2_mul)(&FN(B2), &b, &t);
    }
#else
    FN(fp_from_u64)(FN(XI).c0, 1);
    FN(fp_from_u64)(FN(XI).c1, 1);
    FN(fp_from_u64)(FN(B1), 4);
    FN(fp_from_u64)(FN(B2).c0, 4);
    FN(fp_from_u64)(FN(B2).c1, 4);
#endif
    FN(fp_from_u64)(FN(HALF), 2);
    FN(fp_inv)(FN(HALF), FN(HALF));
    FN(fp2_dbl)(&FN(B2x3), &FN(B2));
    FN(fp2_add)(&FN(B2x3), &FN(B2x3), &FN(B2));
    /* gamma_1 = xi^((p-1)/6) */
    uint64_t e[NL], pm1[NL];
    memcpy(pm1, FN(P), sizeof(pm1));
    pm1[0] -= 1;
    FN(limbs_div_small)(e, pm1, NL, 6);
    memset(&FN(GAMMA)[0], 0, sizeof(fp2));
    FN(fp_set)(FN(GAMMA)[0].c0, FN(R1));
    FN(fp2_pow)(&FN(GAMMA)[1], &FN(XI), e, NL);
    for (int k = 2; k < 6; k++) FN(fp2_mul)(&FN(GAMMA)[k], &FN(GAMMA)[k - 1], &FN(GAMMA)[1]);
    FN(inited) = 1;
}

static void FN(g1_load)(g1aff *P, const uint8_t *in) {
    int allzero = 1;
    for (int i = 0; i < 2 * NL * 8; i++) if (in[i]) { allzero = 0; break; }
    P->inf = allzero || (!BN && (in[0] & 0x40));
    if (P->inf) return;
    FN(fp_from_bytes)(P->x, in);
    FN(fp_from_bytes)(P->y, in + NL * 8);
}
static void FN(g2_load)(g2aff *Q, const uint8_t *in) {
    const int fb = NL * 8;
    int allzero = 1;
    for (int i = 0; i < 4 * fb; i++) if (in[i]) { allzero = 0; break; }
    Q->inf = allzero || (!BN && (in[0] & 0x40));
    if (Q->inf) return;
    FN(fp_from_bytes)(Q->x.c1, in);
    FN(fp_from_bytes)(Q->x.c0, in + fb);
    FN(fp_from_bytes)(Q->y.c1, in + 2 * fb);
    FN(fp_from_bytes)(Q->y.c0, in + 3 * fb);
}

/* product over [lo,hi) of Miller values (mode 0) or of full pairings (mode 1) */
static void FN(pair_range)(fp12 *acc, const uint8_t *g1, const uint8_t *g2, size_t lo, size_t hi, int mode) {
    FN(fp12_one)(acc);
    for (size_t i = lo; i < hi; i++) {
        g1aff P;
        g2aff Q;
        fp12 f;
        FN(g1_load)(&P, g1 + i * 2 * NL * 8);
        FN(g2_load)(&Q, g2 + i * 4 * NL * 8);
        FN(miller)(&f, &P, &Q);
        if (mode == 1) FN(final_exp)(&f, &f);
        FN(fp12_mul)(acc, acc, &f);
    }
}

#undef fp
#undef fp2
#undef fp6
#undef fp12
#undef g2proj
#undef g1aff
#undef g2aff
#undef jac
#undef FE
