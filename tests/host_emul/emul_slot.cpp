// TEST SCAFFOLDING: runs the slot engine's interpreter (bgls_b200/csrc/slotvm.cuh) and its generated programs on the
// host, lane by lane, with the carry primitives emulated (arith.cuh, non-CUDA branch).  Never linked into the product.
#include <cstring>
#include <vector>
#include "../../bgls_b200/csrc/slotvm.cuh"

using namespace bgls;

// Miller product of n pairs as one "block" of NPB groups of K pairs with G lanes per group; out = GT wire record (raw product)
template <class C, class T> static void slot_miller_product(const uint8_t* g1, const uint8_t* g2, int n, uint8_t* out) {
    constexpr int N = C::N, FB = C::FP_BYTES, G = T::G, K = T::K, W4 = 2 * N / 4;
    constexpr int NPB = 8;
    std::vector<SvU4> slots((size_t)T::NSLOT * W4 * NPB), consts((size_t)T::NCONST * W4);
    memcpy(consts.data(), T::consts(), (size_t)T::NCONST * 2 * N * 4);
    // deliberately dirty slot file: programs must not depend on initial contents
    for (auto& s : slots) s = SvU4{0xdeadbeefu, 0x12345678u, 0x9abcdef0u, 0x0badf00du};
    uint32_t inf[NPB];
    for (int q = 0; q < NPB; q++) {
        SlotFile<C, NPB> sf{slots.data(), consts.data(), q, 0};
        uint32_t anyp = 0, anyq = 0, flag = 0;
        for (int t = 0; t < 6 * K; t++) {
            const int j = t / 6, c = t % 6;
            const int pair = q * K + j;
            LN<N> v = sv_zero<C>();
            if (pair < n) {
                const uint8_t* src = c < 2 ? g1 + (size_t)pair * 2 * FB + c * FB : g2 + (size_t)pair * 4 * FB + (c - 2) * FB;
                v = sv_fp_from_be<C>(src);
                uint32_t any = 0;
                for (int i = 0; i < N; i++) any |= v.v[i];
                if (any) { if (c < 2) anyp |= 1u << j; else anyq |= 1u << j; }
                if (!C::IS_BN && (c == 0 || c == 2) && (src[0] & 0x40)) flag |= 1u << j;
            }
            sv_store_coord<C, T, NPB>(sf, j, c, v);
        }
        inf[q] = (~(anyp & anyq) | flag) & ((1u << K) - 1);
        for (int j = 0; j < K; j++) sf.store(T::S_TZ0 + 7 * j, sf.load(SV_CONST0 + 1));
    }
    const uint32_t* code = T::code();
    const uint32_t* offs = T::offsets();
    for (int s = 0; s < T::SEQ_LEN; s++) {
        const uint32_t pid = T::sequence()[s];
        for (uint32_t w = offs[pid]; w < offs[pid + 1]; w += G)
            for (int q = 0; q < NPB; q++) {
                SlotFile<C, NPB> sf{slots.data(), consts.data(), q, inf[q]};
                for (int gl = 0; gl < G; gl++) sv_exec<C, NPB>(sf, code[w + gl]);   // lanes in turn: hazards excluded by the generator
            }
    }
    if (K == 1)
        for (int q = 0; q < NPB; q++) {
            if (!inf[q]) continue;
            SlotFile<C, NPB> sf{slots.data(), consts.data(), q, 0};
            sv_set_one<C, T, NPB>(sf);
        }
    const int ngroups_all = (n + K - 1) / K;
    const int ngroups = ngroups_all < NPB ? ngroups_all : NPB;
    for (int st = 1; st < NPB; st <<= 1) {
        if (st >= ngroups) break;
        for (int q = 0; q < NPB; q++) {
            const bool act = (q & (2 * st - 1)) == 0 && q + st < ngroups;
            if (!act) continue;
            SlotFile<C, NPB> sf{slots.data(), consts.data(), q, 0}, pf{slots.data(), consts.data(), q + st, 0};
            for (int k = 0; k < 6; k++) sf.store(T::S_G00 + k, pf.load(T::S_F00 + k));
            for (uint32_t w = offs[T::P_MUL12]; w < offs[T::P_MUL12 + 1]; w += G)
                for (int gl = 0; gl < G; gl++) sv_exec<C, NPB>(sf, code[w + gl]);
        }
    }
    SlotFile<C, NPB> p0{slots.data(), consts.data(), 0, 0};
    for (int t = 0; t < 12; t++) {
        const int i = t >> 1, part = t & 1;
        sv_fp_to_be<C>(out + (size_t)t * FB, p0.load_fp(sv_wire_slot<T>(i), part ? 0 : 1));
    }
}

extern "C" int emu_slot_miller_product(int curve, int g, int k, const uint8_t* g1, const uint8_t* g2, int n, uint8_t* out) {
    if (n > 8 * k) return -1;
    const int shape = g * 10 + k;
    if (curve == 0) {
        if (shape == 11) slot_miller_product<BN254, svt::BN254_G1>(g1, g2, n, out);
        else if (shape == 21) slot_miller_product<BN254, svt::BN254_G2>(g1, g2, n, out);
        else if (shape == 41) slot_miller_product<BN254, svt::BN254_G4>(g1, g2, n, out);
        else if (shape == 82) slot_miller_product<BN254, svt::BN254_G8K2>(g1, g2, n, out);
        else if (shape == 81) slot_miller_product<BN254, svt::BN254_G8>(g1, g2, n, out);
        else if (shape == 161) slot_miller_product<BN254, svt::BN254_G16>(g1, g2, n, out);
        else if (shape == 42) slot_miller_product<BN254, svt::BN254_G4K2>(g1, g2, n, out);
        else return -1;
    } else {
        if (shape == 11) slot_miller_product<BLS381, svt::BLS381_G1>(g1, g2, n, out);
        else if (shape == 21) slot_miller_product<BLS381, svt::BLS381_G2>(g1, g2, n, out);
        else if (shape == 41) slot_miller_product<BLS381, svt::BLS381_G4>(g1, g2, n, out);
        else if (shape == 82) slot_miller_product<BLS381, svt::BLS381_G8K2>(g1, g2, n, out);
        else if (shape == 81) slot_miller_product<BLS381, svt::BLS381_G8>(g1, g2, n, out);
        else if (shape == 161) slot_miller_product<BLS381, svt::BLS381_G16>(g1, g2, n, out);
        else if (shape == 42) slot_miller_product<BLS381, svt::BLS381_G4K2>(g1, g2, n, out);
        else return -1;
    }
    return 0;
}

// in: a0, a1, b0, b1 (N limbs each); out: fp2 mul (2N), fp2 sqr (2N), xi*a (2N), a0/2 (N), a0+b1 (N), a0-b1 (N), a0*b0 (2N)
template <class C> static void sat_ops(const uint32_t* in, uint32_t* out) {
    constexpr int N = C::N;
    F2<C> a, b;
    for (int i = 0; i < N; i++) { a.c0.v[i] = in[i]; a.c1.v[i] = in[N + i]; b.c0.v[i] = in[2 * N + i]; b.c1.v[i] = in[3 * N + i]; }
    const F2<C> m = sat_fp2_mul<C>(a, b), q = sat_fp2_sqr<C>(a), x = sat_fp2_mul_xi<C>(a);
    LN<N> h;
    mp_half<C>(h.v, a.c0.v);
    const LN<N> s = mp_add_f<C>(a.c0, b.c1), d = mp_sub_f<C>(a.c0, b.c1);
    const LN<2 * N> t = mp_mul_f<N>(a.c0, b.c0);
    for (int i = 0; i < N; i++) {
        out[i] = m.c0.v[i]; out[N + i] = m.c1.v[i]; out[2 * N + i] = q.c0.v[i]; out[3 * N + i] = q.c1.v[i];
        out[4 * N + i] = x.c0.v[i]; out[5 * N + i] = x.c1.v[i]; out[6 * N + i] = h.v[i]; out[7 * N + i] = s.v[i]; out[8 * N + i] = d.v[i];
    }
    for (int i = 0; i < 2 * N; i++) out[9 * N + i] = t.v[i];
}
extern "C" void emu_sat_ops(int curve, const uint32_t* in, uint32_t* out) {
    if (curve == 0) sat_ops<BN254>(in, out); else sat_ops<BLS381>(in, out);
}
