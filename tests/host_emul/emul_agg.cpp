// TEST SCAFFOLDING: runs the six-lane point aggregation (bgls_b200/csrc/agg.cuh) on the host, lane by lane, with the carry
// primitives emulated (arith.cuh, non-CUDA branch).  Never linked into the product.
#include <cstring>
#include <vector>
#include "../../bgls_b200/csrc/agg.cuh"

using namespace bgls;

template <class E, class F> static void add_on_host(const F& f) {
    for (int step = 0; step < 3; step++) {
        for (int r = 0; r < AGG_LANES; r++) {
            if (step == 1 && !E::B3_MUL) agg_b3_small<E>(f, r);
            else if (agg_mul_active<E>(step, r)) {
                typename E::T a, b;
                agg_mul_operands<E>(f, step, r, a, b);
                f.store(agg_mul_dst(step, r), E::mul(a, b));
            }
        }
        for (int r = 0; r < AGG_LANES; r++) agg_linear<E>(f, step, r);
    }
}
// one "block" of NGB groups; group q adds the points q, q + NGB, ...; binary tree; affine record
template <class E> static void aggregate(const uint8_t* pts, size_t n, uint8_t* out) {
    using C = typename E::Curve;
    constexpr int N = C::N, FB = C::FP_BYTES, NGB = 8, NC = 2 * E::HALVES;
    using F = AggFile<E, NGB>;
    std::vector<SvU4> slots((size_t)AG_NSLOT * F::W4 * NGB, SvU4{0xdeadbeefu, 0x12345678u, 0x9abcdef0u, 0x0badf00du});
    for (int q = 0; q < NGB; q++) {
        const F f{slots.data(), q};
        for (int r = 0; r < AGG_LANES; r++) agg_init_acc<E>(f, r);
        for (size_t i = q; i < n; i += NGB) {
            const uint8_t* rec = pts + i * NC * FB;
            LN<N> v[NC];
            bool any = false;
            for (int c = 0; c < NC; c++) { v[c] = agg_read_be<C>(rec + c * FB); any |= !mpw_is_zero<N>(v[c].v); }
            const bool inf = !any || (!C::IS_BN && (rec[0] & 0x40));
            for (int r = 0; r < AGG_LANES; r++) agg_store_input<E>(f, r, r < NC ? v[r] : agg_small<N>(0u), inf);
            add_on_host<E>(f);
        }
    }
    const int ngroups = n < (size_t)NGB ? (n ? (int)n : 1) : NGB;
    // one round trip through the cross-block value format
    std::vector<uint32_t> val(3 * E::HALVES * N);
    for (int st = 1; st < ngroups; st <<= 1)
        for (int q = 0; q + st < ngroups; q += 2 * st) {
            const F f{slots.data(), q}, o{slots.data(), q + st};
            for (int r = 0; r < AGG_LANES; r++) agg_export<E>(o, r, AG_X1, val.data());
            for (int r = 0; r < AGG_LANES; r++) agg_import<E>(f, r, AG_X2, val.data());
            add_on_host<E>(f);
        }
    agg_store_affine<E>(F{slots.data(), 0}, out);
}

extern "C" int emu_agg6(int curve, int group, const uint8_t* pts, size_t n, uint8_t* out) {
    if (curve == 0) { if (group == 1) aggregate<AggFp<BN254>>(pts, n, out); else aggregate<AggFp2<BN254>>(pts, n, out); }
    else { if (group == 1) aggregate<AggFp<BLS381>>(pts, n, out); else aggregate<AggFp2<BLS381>>(pts, n, out); }
    return 0;
}
// plain inverse of a (N limbs, 0 < a < p)
extern "C" int emu_inv_plain(int curve, const uint32_t* a, uint32_t* out) {
    if (curve == 0) { LN<8> x; memcpy(x.v, a, 32); x = mp_inv_plain<BN254>(x); memcpy(out, x.v, 32); }
    else { LN<12> x; memcpy(x.v, a, 48); x = mp_inv_plain<BLS381>(x); memcpy(out, x.v, 48); }
    return 0;
}
// Jacobi symbol (a / p) of the N-limb value a
extern "C" int emu_jacobi(int curve, const uint32_t* a) { return curve == 0 ? mp_jacobi<BN254>(a) : mp_jacobi<BLS381>(a); }
