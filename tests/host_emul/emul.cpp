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
template <class C, int K> static int pairing_product_shared(const uint8_t* g1, const uint8_t* g2, size_t n, uint8_t* out, int do_final) {
    Fp12<C> acc;
    fp12_one(acc);
    for (size_t base = 0; base < n; base += K) {
        G1Aff<C> P[K]; G2Aff<C> Q[K]; Fp12<C> f;
        const int k = (int)(n - base < (size_t)K ? n - base : (size_t)K);
        for (int j = 0; j < k; j++) {
            g1_load<C>(P[j], g1 + (base + j) * 2 * C::FP_BYTES);
            g2_load<C>(Q[j], g2 + (base + j) * 4 * C::FP_BYTES);
        }
        miller_loop_shared<C, K>(f, P, Q, k);
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
int emu_pairing_product_shared(int curve, const uint8_t* g1, const uint8_t* g2, size_t n, uint8_t* out, int do_final) {
    return curve == 0 ? pairing_product_shared<BN254, 4>(g1, g2, n, out, do_final) : pairing_product_shared<BLS381, 4>(g1, g2, n, out, do_final);
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

// ============================================================ dot-product machine (host emulation)
#include <vector>
#include "../../bgls_b200/csrc/machine.cuh"
#include "../../bgls_b200/csrc/machine_tables.cuh"

template <class M, class T> struct HostMach {
    std::vector<uint32_t> gs;
    MachView<M> mv;
    MachTables tb;
    HostMach() : gs((size_t)M::NS * M::L, 0) {
        mach_fill_consts<M>(gs.data(), T::consts());
        mv.gs = gs.data();
        tb.consts = T::consts();
        tb.hdr = T::hdr();
        tb.rec = T::rec();
    }
    void run(const uint16_t* prog, int len) {
        for (int pc = 0; pc < len; pc++) {
            const uint32_t ph = prog[pc];
            const uint32_t h = tb.hdr[ph];
            const int Tn = (h >> 8) & 0xFF;
            uint32_t outs[M::LANES][M::L];
            if constexpr (M::SIGNED) {
                if ((h & 0xFF) == 3) {   // Karatsuba-lane DOT phase: accumulate per lane, combine the (Q, P, S) triples, reduce
                    static unsigned long long acc[M::LANES][2 * M::L];
                    for (int lane = 0; lane < M::LANES; lane++)
                        mach_kdot_acc<M>(mv, tb.rec + ((size_t)ph * M::LANES + lane) * M::REC, Tn, acc[lane]);
                    const int ntr = (h >> 16) & 0xFF;
                    for (int j = 0; j < ntr; j++)
                        for (int c = 0; c < 2 * M::L; c++) {
                            const unsigned long long q = acc[3 * j][c], pp = acc[3 * j + 1][c];
                            acc[3 * j + 1][c] = pp - q;
                            acc[3 * j + 2][c] -= pp + q;
                        }
                    for (int lane = 0; lane < M::LANES; lane++) {
                        const uint16_t* rec = tb.rec + ((size_t)ph * M::LANES + lane) * M::REC;
                        mach_kdot_finish<M>(outs[lane], acc[lane], h >> 24);
                        if (rec[0] != MIDLE) mach_store<M>(mv, rec[0], outs[lane]);
                    }
                    continue;
                }
            }
            for (int lane = 0; lane < M::LANES; lane++) {
                const uint16_t* rec = tb.rec + ((size_t)ph * M::LANES + lane) * M::REC;
                if ((h & 0xFF) == 2) {
                    if (rec[0] != MIDLE) { uint32_t in[M::L]; mach_load<M>(in, mv, rec[1]); mach_inv<M>(outs[lane], in); }
                } else if ((h & 0xFF) == 0) mach_dot<M>(mv, rec, Tn, h >> 24, outs[lane]);
                else mach_lin<M>(mv, rec, Tn, outs[lane]);
            }
            for (int lane = 0; lane < M::LANES; lane++) {  // deferred stores: a phase must be hazard-free
                const uint16_t* rec = tb.rec + ((size_t)ph * M::LANES + lane) * M::REC;
                if (rec[0] != MIDLE) mach_store<M>(mv, rec[0], outs[lane]);
            }
        }
    }
    void set(int s, const uint32_t* v) { mach_store<M>(mv, s, v); }
    void get(int s, uint32_t* v) { mach_load<M>(v, mv, s); }
};

// PM/PT: slot file that runs the Miller program (the 16-lane M file or the pipelined 32-lane P file)
template <class MM, class MT, class FM, class FT, class PM = MM, class PT = MT>
static int mach_pairing(const uint8_t* g1, const uint8_t* g2, size_t n, uint8_t* out, int do_final, uint32_t* dbg_f) {
    constexpr int L = MM::L, FB = MM::FP_BYTES;
    static_assert(PM::L == MM::L && PM::W == MM::W, "limb layouts must agree");
    HostMach<MM, MT> acc;   // accumulator group: product tree through MUL_AB / MUL_BA
    HostMach<PM, PT> accp;  // P file: in-place MULACC (the block trees of k_mach_miller32 / k_mach_tree32)
    bool have = false, in_a = true;
    uint32_t one[L], zero[L];
    acc.get(MM::ONE, one);
    acc.get(MM::ZERO, zero);
    for (size_t i = 0; i < n; i++) {
        HostMach<PM, PT> m;
        const uint8_t* p1 = g1 + i * 2 * FB;
        const uint8_t* p2 = g2 + i * 4 * FB;
        bool inf = bytes_all_zero(p1, 2 * FB) || bytes_all_zero(p2, 4 * FB) || (FB == 48 && ((p1[0] & 0x40) || (p2[0] & 0x40)));
        uint32_t f[12][L];
        if (inf) {
            for (int k = 0; k < 12; k++) memcpy(f[k], k == 0 ? one : zero, sizeof(one));
        } else {
            uint32_t v[L];
            mach_limbs_from_be<PM>(v, p1); m.set(PM::IN_XP, v);
            mach_limbs_from_be<PM>(v, p1 + FB); m.set(PM::IN_YP, v);
            mach_limbs_from_be<PM>(v, p2); m.set(PM::IN_XQY, v);
            mach_limbs_from_be<PM>(v, p2 + FB); m.set(PM::IN_XQX, v);
            mach_limbs_from_be<PM>(v, p2 + 2 * FB); m.set(PM::IN_YQY, v);
            mach_limbs_from_be<PM>(v, p2 + 3 * FB); m.set(PM::IN_YQX, v);
            m.run(PT::prog_MILLER(), PM::MILLER_LEN);
            for (int k = 0; k < 12; k++) m.get(PM::FA0 + k, f[k]);
        }
        if (dbg_f && i == 0) memcpy(dbg_f, f, sizeof(f));
        if constexpr (PM::SIGNED) {
            if (!have) {
                for (int k = 0; k < 12; k++) accp.set(PM::FA0 + k, f[k]);
                have = true;
            } else {
                for (int k = 0; k < 12; k++) accp.set(PM::GB0 + k, f[k]);
                accp.run(PT::prog_MULACC(), PM::MULACC_LEN);
            }
        } else if (!have) {
            for (int k = 0; k < 12; k++) acc.set(MM::FA0 + k, f[k]);
            have = true; in_a = true;
        } else {
            for (int k = 0; k < 12; k++) acc.set(MM::GB0 + k, f[k]);
            if (in_a) acc.run(MT::prog_MUL_AB(), MM::MUL_AB_LEN); else acc.run(MT::prog_MUL_BA(), MM::MUL_BA_LEN);
            in_a = !in_a;
        }
    }
    uint32_t f[12][L];
    for (int k = 0; k < 12; k++) {
        if (!have) memcpy(f[k], k == 0 ? one : zero, sizeof(one));
        else if constexpr (PM::SIGNED) accp.get(PM::FA0 + k, f[k]);
        else acc.get((in_a ? MM::FA0 : MM::FB0) + k, f[k]);
    }
    HostMach<FM, FT> fe;
    for (int k = 0; k < 12; k++) fe.set(FM::FA0 + k, f[k]);
    if (do_final) {
        fe.run(FT::prog_FINALEXP(), FM::FINALEXP_LEN);
    } else {
        fe.run(FT::prog_EXPORT(), FM::EXPORT_LEN);
    }
    // OUT slots: (k, re/im) -> GT layout: w-powers 5,3,1,4,2,0 each (im, re)
    const int order[6] = {5, 3, 1, 4, 2, 0};
    bool all_one = true;
    for (int i = 0; i < 6; i++) {
        for (int c = 0; c < 2; c++) {  // c = 0: im first
            uint32_t v[L];
            fe.get(FM::OUT0 + 2 * order[i] + (c == 0 ? 1 : 0), v);
            bool z, o;
            mach_canon_be<FM>(out + (2 * i + c) * FB, v, &z, &o);
            const bool want_one = (order[i] == 0 && c == 1);
            all_one = all_one && (want_one ? o : z);
        }
    }
    return all_one ? 1 : 0;
}

extern "C" int emu_mach_pairing_product(int curve, const uint8_t* g1, const uint8_t* g2, size_t n, uint8_t* out, int do_final, uint32_t* dbg_f) {
    using namespace mtab;
    if (curve == 0) return mach_pairing<BN254_M, BN254_M_T, BN254_F, BN254_F_T>(g1, g2, n, out, do_final, dbg_f);
    return mach_pairing<BLS381_M, BLS381_M_T, BLS381_F, BLS381_F_T>(g1, g2, n, out, do_final, dbg_f);
}
// same pipeline with the Miller loops run by the pipelined 32-lane program (slot file P)
extern "C" int emu_mach_pairing_product_p(int curve, const uint8_t* g1, const uint8_t* g2, size_t n, uint8_t* out, int do_final) {
    using namespace mtab;
    if (curve == 0) return mach_pairing<BN254_M, BN254_M_T, BN254_F, BN254_F_T, BN254_MP, BN254_MP_T>(g1, g2, n, out, do_final, nullptr);
    return mach_pairing<BLS381_M, BLS381_M_T, BLS381_F, BLS381_F_T, BLS381_MP, BLS381_MP_T>(g1, g2, n, out, do_final, nullptr);
}

// ============================================================ hash-to-G1 (host emulation of the device functions)
#include "../../bgls_b200/csrc/hash.cuh"
extern "C" void emu_hash_to_g1(int curve, const uint8_t* msg, size_t len, uint8_t* out) {
    if (curve == 0) hash_to_g1_keccak_ti<BN254>(out, msg, len);
    else hash_to_g1_ft<BLS381>(out, msg, len);
}

// ============================================================ compressed wire formats (host emulation of codec.cuh)
#include "../../bgls_b200/csrc/codec.cuh"
extern "C" void emu_compress(int curve, int group, const uint8_t* rec, uint8_t* out) {
    if (curve == 0) { if (group == 1) compress_g1<BN254>(out, rec); else compress_g2<BN254>(out, rec); }
    else { if (group == 1) compress_g1<BLS381>(out, rec); else compress_g2<BLS381>(out, rec); }
}
extern "C" int emu_decompress(int curve, int group, const uint8_t* in, int check_subgroup, uint8_t* rec) {
    if (curve == 0) return group == 1 ? decompress_g1<BN254>(rec, in, check_subgroup) : decompress_g2<BN254>(rec, in, check_subgroup);
    return group == 1 ? decompress_g1<BLS381>(rec, in, check_subgroup) : decompress_g2<BLS381>(rec, in, check_subgroup);
}
// throughput form of the bls12-381 hash (one cofactor multiplication for both halves)
extern "C" void emu_hash_to_g1_shared_cofactor(const uint8_t* msg, size_t len, uint8_t* out) { hash_to_g1_ft_shared_cofactor<BLS381>(out, msg, len); }
