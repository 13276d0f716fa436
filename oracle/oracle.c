/* TEST INFRASTRUCTURE (oracle) -- not part of the product path.
 *
 * C restatement of the Project-Arda/bgls hot path (CPU, exact).  Built by oracle/Makefile into
 * oracle/liboracle.so; loaded only by tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs.
 *
 * Boundary formats are the reference's uncompressed affine big-endian records
 * (curves/altbn128.go:149-158, curves/bls12_381.go:147-158): G1 x||y, G2 x_im||x_re||y_im||y_re,
 * infinity = all-zero record (bls12 also accepts the 0x40 flag byte).
 *
 * orc_pairing_product mode 1 mirrors concurrentPairingProduct (curves/curve.go:125-170):
 * every pair pays a full Pair() = Miller loop + final exponentiation, fanned out over threads,
 * then the GT values are multiplied.  mode 0 multiplies Miller values and exponentiates once.
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef unsigned __int128 u128;
static __thread uint64_t orc_fpmul_counter = 0;

#define PFX bn_
#define NL 4
#define BN 1
#include "pairing_impl.h"
#undef PFX
#undef NL
#undef BN

#define PFX bl_
#define NL 6
#define BN 0
#include "pairing_impl.h"
#undef PFX
#undef NL
#undef BN

#define ORC_ALTBN128 0
#define ORC_BLS12 1

static pthread_once_t once = PTHREAD_ONCE_INIT;
static void init_all(void) {
    bn_init();
    bl_init();
}
static void ensure_init(void) { pthread_once(&once, init_all); }

uint64_t orc_fpmul_count(void) { return orc_fpmul_counter; }
void orc_fpmul_reset(void) { orc_fpmul_counter = 0; }

/* ------------------------------------------------------------------ threaded range helper */
typedef struct {
    int curve, mode;
    const uint8_t *g1, *g2;
    size_t lo, hi;
    union { bn_fp12 bn; bl_fp12 bl; } acc;
} pair_job;

static void *pair_worker(void *arg) {
    pair_job *j = (pair_job *)arg;
    if (j->curve == ORC_ALTBN128) bn_pair_range(&j->acc.bn, j->g1, j->g2, j->lo, j->hi, j->mode);
    else bl_pair_range(&j->acc.bl, j->g1, j->g2, j->lo, j->hi, j->mode);
    return NULL;
}

/* out: 12 Fp values in GT layout.  do_final: 0 = raw Miller product, 1 = exponentiate. */
static int pairing_product_impl(int curve, const uint8_t *g1, const uint8_t *g2, size_t n, uint8_t *out,
                                int nthreads, int mode, int do_final) {
    ensure_init();
    if (curve != ORC_ALTBN128 && curve != ORC_BLS12) return -1;
    if (nthreads < 1) nthreads = 1;
    if ((size_t)nthreads > n) nthreads = n ? (int)n : 1;
    pair_job *jobs = (pair_job *)calloc(nthreads, sizeof(pair_job));
    pthread_t *th = (pthread_t *)calloc(nthreads, sizeof(pthread_t));
    for (int t = 0; t < nthreads; t++) {
        jobs[t].curve = curve;
        jobs[t].mode = mode;
        jobs[t].g1 = g1;
        jobs[t].g2 = g2;
        jobs[t].lo = n * t / nthreads;
        jobs[t].hi = n * (t + 1) / nthreads;
        if (nthreads > 1) pthread_create(&th[t], NULL, pair_worker, &jobs[t]);
        else pair_worker(&jobs[t]);
    }
    if (nthreads > 1)
        for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
    if (curve == ORC_ALTBN128) {
        bn_fp12 acc = jobs[0].acc.bn;
        for (int t = 1; t < nthreads; t++) bn_fp12_mul(&acc, &acc, &jobs[t].acc.bn);
        if (mode == 0 && do_final) bn_final_exp(&acc, &acc);
        bn_fp12_to_bytes(out, &acc);
    } else {
        bl_fp12 acc = jobs[0].acc.bl;
        for (int t = 1; t < nthreads; t++) bl_fp12_mul(&acc, &acc, &jobs[t].acc.bl);
        if (mode == 0 && do_final) bl_final_exp(&acc, &acc);
        bl_fp12_to_bytes(out, &acc);
    }
    free(jobs);
    free(th);
    return 0;
}

int orc_pairing_product(int curve, const uint8_t *g1, const uint8_t *g2, size_t n, uint8_t *out_gt, int nthreads, int mode) {
    return pairing_product_impl(curve, g1, g2, n, out_gt, nthreads, mode, 1);
}
int orc_miller_product(int curve, const uint8_t *g1, const uint8_t *g2, size_t n, uint8_t *out_f, int nthreads) {
    return pairing_product_impl(curve, g1, g2, n, out_f, nthreads, 0, 0);
}
/* multiply k Fp12 values (GT layout) and apply the final exponentiation (do_final) */
int orc_fp12_product(int curve, const uint8_t *in, size_t k, int do_final, uint8_t *out) {
    ensure_init();
    if (curve == ORC_ALTBN128) {
        bn_fp12 acc, t;
        bn_fp12_one(&acc);
        for (size_t i = 0; i < k; i++) { bn_fp12_from_bytes(&t, in + i * 384); bn_fp12_mul(&acc, &acc, &t); }
        if (do_final) bn_final_exp(&acc, &acc);
        bn_fp12_to_bytes(out, &acc);
    } else if (curve == ORC_BLS12) {
        bl_fp12 acc, t;
        bl_fp12_one(&acc);
        for (size_t i = 0; i < k; i++) { bl_fp12_from_bytes(&t, in + i * 576); bl_fp12_mul(&acc, &acc, &t); }
        if (do_final) bl_final_exp(&acc, &acc);
        bl_fp12_to_bytes(out, &acc);
    } else return -1;
    return 0;
}
/* f^e with e given as little-endian 64-bit limbs: lets python check final_exp against the definition */
int orc_fp12_pow(int curve, const uint8_t *in, const uint64_t *e, int nlimbs, uint8_t *out) {
    ensure_init();
    if (curve == ORC_ALTBN128) {
        bn_fp12 t;
        bn_fp12_from_bytes(&t, in);
        bn_final_exp_slow(&t, &t, e, nlimbs);
        bn_fp12_to_bytes(out, &t);
    } else if (curve == ORC_BLS12) {
        bl_fp12 t;
        bl_fp12_from_bytes(&t, in);
        bl_final_exp_slow(&t, &t, e, nlimbs);
        bl_fp12_to_bytes(out, &t);
    } else return -1;
    return 0;
}

/* ------------------------------------------------------------------ point aggregation / scaling */
typedef struct {
    int curve, k, op; /* op 0: sum range into acc; op 1: scalar-mul each point */
    const uint8_t *pts, *scalars;
    uint8_t *out;
    size_t lo, hi;
    union { bn_jac bn; bl_jac bl; } acc;
} pt_job;

static void *pt_worker(void *arg) {
    pt_job *j = (pt_job *)arg;
    const int k = j->k;
    if (j->curve == ORC_ALTBN128) {
        const size_t rec = (size_t)2 * k * 32;
        if (j->op == 0) {
            memset(&j->acc.bn, 0, sizeof(j->acc.bn));
            j->acc.bn.inf = 1;
            for (size_t i = j->lo; i < j->hi; i++) {
                bn_jac p;
                bn_jac_from_bytes(k, &p, j->pts + i * rec);
                bn_jac_add(k, &j->acc.bn, &j->acc.bn, &p);
            }
        } else {
            for (size_t i = j->lo; i < j->hi; i++) {
                bn_jac p, r;
                bn_jac_from_bytes(k, &p, j->pts + i * rec);
                bn_jac_mul(k, &r, &p, j->scalars + i * 32);
                bn_jac_to_bytes(k, j->out + i * rec, &r);
            }
        }
    } else {
        const size_t rec = (size_t)2 * k * 48;
        if (j->op == 0) {
            memset(&j->acc.bl, 0, sizeof(j->acc.bl));
            j->acc.bl.inf = 1;
            for (size_t i = j->lo; i < j->hi; i++) {
                bl_jac p;
                bl_jac_from_bytes(k, &p, j->pts + i * rec);
                bl_jac_add(k, &j->acc.bl, &j->acc.bl, &p);
            }
        } else {
            for (size_t i = j->lo; i < j->hi; i++) {
                bl_jac p, r;
                bl_jac_from_bytes(k, &p, j->pts + i * rec);
                bl_jac_mul(k, &r, &p, j->scalars + i * 32);
                bl_jac_to_bytes(k, j->out + i * rec, &r);
            }
        }
    }
    return NULL;
}

static int pt_run(int curve, int group, int op, const uint8_t *pts, const uint8_t *scalars, size_t n, uint8_t *out, int nthreads) {
    ensure_init();
    if ((curve != ORC_ALTBN128 && curve != ORC_BLS12) || (group != 1 && group != 2)) return -1;
    if (nthreads < 1) nthreads = 1;
    if ((size_t)nthreads > n) nthreads = n ? (int)n : 1;
    pt_job *jobs = (pt_job *)calloc(nthreads, sizeof(pt_job));
    pthread_t *th = (pthread_t *)calloc(nthreads, sizeof(pthread_t));
    for (int t = 0; t < nthreads; t++) {
        jobs[t].curve = curve; jobs[t].k = group; jobs[t].op = op;
        jobs[t].pts = pts; jobs[t].scalars = scalars; jobs[t].out = out;
        jobs[t].lo = n * t / nthreads; jobs[t].hi = n * (t + 1) / nthreads;
        if (nthreads > 1) pthread_create(&th[t], NULL, pt_worker, &jobs[t]);
        else pt_worker(&jobs[t]);
    }
    if (nthreads > 1) for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
    if (op == 0) {
        if (curve == ORC_ALTBN128) {
            bn_jac acc = jobs[0].acc.bn;
            for (int t = 1; t < nthreads; t++) bn_jac_add(group, &acc, &acc, &jobs[t].acc.bn);
            bn_jac_to_bytes(group, out, &acc);
        } else {
            bl_jac acc = jobs[0].acc.bl;
            for (int t = 1; t < nthreads; t++) bl_jac_add(group, &acc, &acc, &jobs[t].acc.bl);
            bl_jac_to_bytes(group, out, &acc);
        }
    }
    free(jobs); free(th);
    return 0;
}
/* AggregatePoints (curves/curve.go:73-110): sum of n points; group 1 = G1, 2 = G2 */
int orc_aggregate(int curve, int group, const uint8_t *pts, size_t n, uint8_t *out, int nthreads) {
    return pt_run(curve, group, 0, pts, NULL, n, out, nthreads);
}
/* ScalePoints (curves/curve.go:190-214): out[i] = scalars[i] * pts[i]; scalars are 32-byte big-endian */
int orc_scale_points(int curve, int group, const uint8_t *pts, const uint8_t *scalars, size_t n, uint8_t *out, int nthreads) {
    return pt_run(curve, group, 1, pts, scalars, n, out, nthreads);
}
int orc_on_curve(int curve, int group, const uint8_t *pt) {
    ensure_init();
    if (curve == ORC_ALTBN128) { bn_jac p; bn_jac_from_bytes(group, &p, pt); return bn_on_curve(group, &p); }
    bl_jac p; bl_jac_from_bytes(group, &p, pt); return bl_on_curve(group, &p);
}
