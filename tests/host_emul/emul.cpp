// TEST SCAFFOLDING: compiles the device arithmetic templates (bgls_b200/csrc/*.cuh) for the
// host with the carry-chain primitives emulated (arith.cuh, non-CUDA branch), so the exact
// algorithms the kernels run can be checked against the oracle on a CPU-only machine.
// Never linked into the product library.
#include <cstring>
#include "../../bgls_b200/csrc/pairing.cuh"

using namespace bgls;

template <class C> static int pairing_product(const uint8_t* g1, const uint8_t* g2, size_t n, uint8_t* out, int do_final) {
    Fp12<C> acc;
    fp12_one(acc);
    for (size_t i = 0; i < n; i++) {
        G1Aff<C> P; G2Aff<C> Q; Fp12<C> f;
        g1_load<C>(P, g1 + i * 2 * C::FP_BYTES);
        g2_load<C>(Q, g2 + i * 4 * C::FP_BYTES);
        miller_loop(f, P, Q);
        fp12_mul(acc, acc, f);
    }
    if (do_final) final_exp(acc, acc);
    fp12_to_be<C>(out, acc);
    return fp12_is_one(acc) ? 1 : 0;
}
template <class C, class F> static void aggregate(const uint8_t* pts, size_t n, size_t rec, uint8_t* out) {
    Jac<F> acc; acc.inf = true;
    for (size_t i = 0; i < n; i++) { Jac<F> p; jac_load<C>(p, pts + i * rec); jac_add(acc, acc, p); }
    jac_store<C>(out, acc);
}
template <class C, class F> static void scale(const uint8_t* pts, const uint8_t* sc, size_t n, size_t rec, uint8_t* out) {
    for (size_t i = 0; i < n; i++) { Jac<F> p, r; jac_load<C>(p, pts + i * rec); jac_mul(r, p, sc + 32 * i); jac_store<C>(out + i * rec, r); }
}
template <class C> static void fpmul(const uint8_t* a, const uint8_t* b, uint8_t* out) {
    Fp<C> x, y, z;
    fp_from_be<C>(x, a); fp_from_be<C>(y, b);
    fp_mul(z, x, y);
    fp_to_be<C>(out, z);
    Fp<C> s; fp_add(s, x, y); fp_to_be<C>(out + C::FP_BYTES, s);
    fp_sub(s, x, y); fp_to_be<C>(out + 2 * C::FP_BYTES, s);
    fp_inv(s, x); fp_to_be<C>(out + 3 * C::FP_BYTES, s);
}

extern "C" {
int emu_pairing_product(int curve, const uint8_t* g1, const uint8_t* g2, size_t n, uint8_t* out, int do_final) {
    return curve == 0 ? pairing_product<BN254>(g1, g2, n, out, do_final) : pairing_product<BLS381>(g1, g2, n, out, do_final);
}
void emu_aggregate(int curve, int group, const uint8_t* pts, size_t n, uint8_t* out) {
    if (curve == 0) { if (group == 1) aggregate<BN254, Fp<BN254>>(pts, n, 64, out); else aggregate<BN254, Fp2<BN254>>(pts, n, 128, out); }
    else { if (group == 1) aggregate<BLS381, Fp<BLS381>>(pts, n, 96, out); else aggregate<BLS381, Fp2<BLS381>>(pts, n, 192, out); }
}
void emu_scale(int curve, int group, const uint8_t* pts, const uint8_t* sc, size_t n, uint8_t* out) {
    if (curve == 0) { if (group == 1) scale<BN254, Fp<BN254>>(pts, sc, n, 64, out); else scale<BN254, Fp2<BN254>>(pts, sc, n, 128, out); }
    else { if (group == 1) scale<BLS381, Fp<BLS381>>(pts, sc, n, 96, out); else scale<BLS381, Fp2<BLS381>>(pts, sc, n, 192, out); }
}
void emu_fpops(int curve, const uint8_t* a, const uint8_t* b, uint8_t* out) {
    if (curve == 0) fpmul<BN254>(a, b, out); else fpmul<BLS381>(a, b, out);
}
}
