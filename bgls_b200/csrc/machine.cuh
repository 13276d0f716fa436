// Interpreter of the warp-cooperative "dot-product machine" (schedules: tools/gen_machine.py,
// tables: machine_tables.cuh).
//
// A group of 16 or 32 lanes owns one pairing.  All state is a file of Fp slots in shared memory, stored
// word-major (limb i of slot s at gs[i*NS + s]) so that lanes reading different slots hit different
// banks; the file holds the NSG group slots followed by a private copy of the NCONST constants, so
// every operand is addressed the same way (one base register + immediate offsets).
// Signed slot files (M::SIGNED): bit 15 of a DOT term's first operand negates that operand and the
// columns are signed; the phase header carries K so that K p is added to the reduced value (never negative).  Fp elements are L limbs of 28 bits (unsaturated), Montgomery form with
// R = 2^(28 L); values are only kept bounded (bounds proven statically by the generator).
//   DOT task:  dst = MontRed( sum_t a_t * b_t )   -- products accumulate carry-free in 64-bit columns
//              with full-rate IMAD.WIDE.U32 (measured 2x the throughput of the carry-chained form),
//              one lazy reduction per output
//   LIN task:  dst = Normalize( sum_t c_t * s_t  |  |c_t| * (KP - s_t) )
// Host build (tests/host_emul) runs the same code sequentially over the lanes.
#pragma once
#include <cstdint>

#include "arith.cuh"

namespace bgls {

constexpr int MG = 16;        // lanes per group of the Miller / product slot files (two groups per warp)
constexpr uint32_t MIDLE = 0xFFFFu;
// per slot file (generated struct M): LANES lanes per group, lane record = REC u16:
//   [0] dst, [1..TM] a slots (LIN: sources), [1+TM..2TM] b slots (LIN: coefficient words)

template <class M> struct MachView {
    uint32_t* gs;        // group slot file, word-major, NS = NSG + NCONST slots
};

template <class M> HD void mach_load(uint32_t* v, const MachView<M>& mv, uint32_t s) {
    const uint32_t* base = mv.gs + s;
#pragma unroll
    for (int i = 0; i < M::L; i++) v[i] = base[i * M::NS];
}
template <class M> HD void mach_store(const MachView<M>& mv, uint32_t s, const uint32_t* v) {
#pragma unroll
    for (int i = 0; i < M::L; i++) mv.gs[i * M::NS + s] = v[i];
}
// copies the constants of the slot file into the tail of one group file
template <class M> HD void mach_fill_consts(uint32_t* gs, const uint32_t* consts /* [NCONST][L] */, int first = 0, int step = 1) {
    for (int idx = first; idx < M::NCONST * M::L; idx += step) {
        const int c = idx / M::L, i = idx % M::L;
        gs[i * M::NS + M::NSG + c] = consts[idx];
    }
}

// Montgomery reduction of 2L 64-bit columns (radix 2^W) to L normalised limbs: (T + m p) / R
template <class M, class ACC> HD void mach_montred(uint32_t* out, ACC* acc) {
    constexpr int L = M::L, W = M::W;
    constexpr uint32_t MASK = (1u << W) - 1;
#pragma unroll
    for (int i = 0; i < L; i++) {
        const uint32_t m = ((uint32_t)acc[i] * M::N0) & MASK;
#pragma unroll
        for (int j = 0; j < L; j++) acc[i + j] += (ACC)((unsigned long long)m * M::p(j));
        acc[i + 1] += acc[i] >> W;   // exact: the low W bits are zero (arithmetic shift for signed columns)
    }
#pragma unroll
    for (int k = L; k < 2 * L - 1; k++) {
        acc[k + 1] += acc[k] >> W;
        out[k - L] = (uint32_t)acc[k] & MASK;
    }
    out[L - 1] = (uint32_t)acc[2 * L - 1];
}

// one DOT task; rec: lane record, T: number of terms of the phase (uniform over the warp).
// The operands of term t+1 are fetched while the L*L multiply-accumulates of term t issue.
template <class M> HD void mach_dot(const MachView<M>& mv, const uint16_t* rec, int T, uint32_t K, uint32_t* out) {
    constexpr int L = M::L;
    if constexpr (M::SIGNED) {
        long long acc[2 * L];
#pragma unroll
        for (int i = 0; i < L; i++) { acc[i] = 0; acc[L + i] = (long long)((unsigned long long)K * M::p(i)); }
        uint32_t a[L], b[L];
        uint32_t ra = rec[1];
        mach_load<M>(a, mv, ra & 0x7FFFu);
        mach_load<M>(b, mv, rec[1 + M::TM]);
        for (int t = 0; t < T; t++) {
            uint32_t na[L], nb[L];
            const int tn = t + 1 < T ? t + 1 : t;
            const uint32_t rn = rec[1 + tn];
            mach_load<M>(na, mv, rn & 0x7FFFu);
            mach_load<M>(nb, mv, rec[1 + M::TM + tn]);
            const bool neg = (ra & 0x8000u) != 0;
            int sa[L];
#pragma unroll
            for (int i = 0; i < L; i++) sa[i] = neg ? -(int)a[i] : (int)a[i];
#pragma unroll
            for (int i = 0; i < L; i++)
#pragma unroll
                for (int j = 0; j < L; j++) acc[i + j] += (long long)sa[i] * (long long)(int)b[j];
#pragma unroll
            for (int i = 0; i < L; i++) { a[i] = na[i]; b[i] = nb[i]; }
            ra = rn;
        }
        mach_montred<M, long long>(out, acc);
    } else {
        unsigned long long acc[2 * L];
#pragma unroll
        for (int i = 0; i < 2 * L; i++) acc[i] = 0;
        uint32_t a[L], b[L];
        mach_load<M>(a, mv, rec[1]);
        mach_load<M>(b, mv, rec[1 + M::TM]);
        for (int t = 0; t < T; t++) {
            uint32_t na[L], nb[L];
            const int tn = t + 1 < T ? t + 1 : t;
            mach_load<M>(na, mv, rec[1 + tn]);
            mach_load<M>(nb, mv, rec[1 + M::TM + tn]);
#pragma unroll
            for (int i = 0; i < L; i++)
#pragma unroll
                for (int j = 0; j < L; j++) acc[i + j] += (unsigned long long)a[i] * b[j];
#pragma unroll
            for (int i = 0; i < L; i++) { a[i] = na[i]; b[i] = nb[i]; }
        }
        mach_montred<M, unsigned long long>(out, acc);
    }
}

// Karatsuba-lane DOT phase (kind 3, signed files): every term is (a1 + a2) * (b1 + b2) with the operand sums formed
// on the fly (W+1 bits).  Lanes 3j, 3j+1, 3j+2 accumulate the Q = sum a.y b.y, P = sum a.x b.x,
// S = sum (a.x+a.y)(b.x+b.y) parts of one Fp2 sum of products modulo 2^64 (unsigned: S may wrap, exactly); after the
// combination (mach_kdot_combine on the device, by warp shuffle) the P lane reduces re = P - Q and the S lane
// im = S - P - Q, which the generator proves to fit the signed range (Gen.kdot_fits).
// record: [0] dst, [1 + 2t], [2 + 2t] = a1 (bit 15: minus), a2, [1 + TM + 2t], [2 + TM + 2t] = b1, b2
template <class M> HD void mach_kdot_acc(const MachView<M>& mv, const uint16_t* rec, int T, unsigned long long* acc) {
    constexpr int L = M::L;
#pragma unroll
    for (int i = 0; i < 2 * L; i++) acc[i] = 0;
    for (int t = 0; t < T; t++) {
        const uint32_t ra = rec[1 + 2 * t];
        uint32_t a1[L], a2[L], b1[L], b2[L];
        mach_load<M>(a1, mv, ra & 0x7FFFu);
        mach_load<M>(a2, mv, rec[2 + 2 * t]);
        mach_load<M>(b1, mv, rec[1 + M::TM + 2 * t]);
        mach_load<M>(b2, mv, rec[2 + M::TM + 2 * t]);
        const bool neg = (ra & 0x8000u) != 0;
        int sa[L], sb[L];
#pragma unroll
        for (int i = 0; i < L; i++) {
            const int x = (int)(a1[i] + a2[i]);
            sa[i] = neg ? -x : x;
            sb[i] = (int)(b1[i] + b2[i]);
        }
#pragma unroll
        for (int i = 0; i < L; i++)
#pragma unroll
            for (int j = 0; j < L; j++) acc[i + j] += (unsigned long long)((long long)sa[i] * (long long)sb[j]);
    }
}
template <class M> HD void mach_kdot_finish(uint32_t* out, const unsigned long long* uacc, uint32_t K) {
    constexpr int L = M::L;
    long long acc[2 * L];
#pragma unroll
    for (int i = 0; i < L; i++) {
        acc[i] = (long long)uacc[i];
        acc[L + i] = (long long)(uacc[L + i] + (unsigned long long)K * M::p(i));
    }
    mach_montred<M, long long>(out, acc);
}
#ifdef __CUDA_ARCH__
// lanes (3j, 3j+1, 3j+2) = (Q, P, S) for j < ntr: P <- P - Q, S <- S - (P + Q); two shuffles per column
template <class M> __device__ __forceinline__ void mach_kdot_combine(unsigned long long* acc, int ntr, int lane) {
    const int role = lane < 3 * ntr ? lane % 3 : -1;
#pragma unroll
    for (int c = 0; c < 2 * M::L; c++) {
        const unsigned long long x = __shfl_up_sync(0xFFFFFFFFu, acc[c], 1);
        unsigned long long sum = acc[c];
        if (role == 1) { sum = acc[c] + x; acc[c] -= x; }
        const unsigned long long y = __shfl_up_sync(0xFFFFFFFFu, sum, 1);
        if (role == 2) acc[c] -= y;
    }
}
#endif

template <class M> HD void mach_lin(const MachView<M>& mv, const uint16_t* rec, int T, uint32_t* out) {
    constexpr int L = M::L, W = M::W;
    constexpr uint32_t MASK = (1u << W) - 1;
    if constexpr (M::SIGNED) {
        // signed files: dst = Normalize(sum c_t * s_t + k p) with two's complement coefficients and ONE offset constant
        // per task (record field 2 TM + 1), signed columns, arithmetic carries
        long long acc[L];
        uint32_t k[L];
        mach_load<M>(k, mv, rec[2 * M::TM + 1]);
#pragma unroll
        for (int i = 0; i < L; i++) acc[i] = (long long)k[i];
        for (int t = 0; t < T; t++) {
            const int c = (int)(signed char)(rec[1 + M::TM + t] & 0xFFu);
            uint32_t v[L];
            mach_load<M>(v, mv, rec[1 + t]);
#pragma unroll
            for (int i = 0; i < L; i++) acc[i] += (long long)c * (long long)(int)v[i];
        }
#pragma unroll
        for (int i = 0; i < L - 1; i++) {
            acc[i + 1] += acc[i] >> W;
            out[i] = (uint32_t)acc[i] & MASK;
        }
        out[L - 1] = (uint32_t)acc[L - 1];
        return;
    }
    unsigned long long acc[L];
#pragma unroll
    for (int i = 0; i < L; i++) acc[i] = 0;
    for (int t = 0; t < T; t++) {
        const uint32_t cw = rec[1 + M::TM + t];
        const uint32_t c = cw & 0x7Fu;
        const bool neg = (cw & 0x80u) != 0;
        uint32_t v[L], k[L];
        mach_load<M>(v, mv, rec[1 + t]);
        mach_load<M>(k, mv, M::NSG + (cw >> 8));  // KP constant (slot 0 of the constants when unused)
#pragma unroll
        for (int i = 0; i < L; i++) {
            const uint32_t x = neg ? k[i] - v[i] : v[i];
            acc[i] += (unsigned long long)c * x;
        }
    }
#pragma unroll
    for (int i = 0; i < L - 1; i++) {
        acc[i + 1] += acc[i] >> W;
        out[i] = (uint32_t)acc[i] & MASK;
    }
    out[L - 1] = (uint32_t)acc[L - 1];
}

// big-endian field element (FP_BYTES) -> L raw limbs of W bits
template <class M> HD void mach_limbs_from_be(uint32_t* out, const uint8_t* be) {
    constexpr int NW = M::FP_BYTES / 4, L = M::L, W = M::W;
    constexpr uint32_t MASK = (1u << W) - 1;
    uint32_t w[NW + 1];
#pragma unroll
    for (int i = 0; i < NW; i++) {
        const uint8_t* q = be + 4 * (NW - 1 - i);
        w[i] = ((uint32_t)q[0] << 24) | ((uint32_t)q[1] << 16) | ((uint32_t)q[2] << 8) | q[3];
    }
    w[NW] = 0;
#pragma unroll
    for (int i = 0; i < L; i++) {
        const int bit = W * i, wi = bit >> 5, sh = bit & 31;
        static_assert(W * (L - 1) / 32 < NW, "top limb must start inside the input words");
        const uint32_t lo = w[wi], hi = w[wi + 1];  // w[NW] == 0
        const uint32_t x = sh ? ((lo >> sh) | (hi << (32 - sh))) : lo;
        out[i] = x & MASK;
    }
}
// L limbs holding a value < 2p (plain domain) -> canonical big-endian bytes; returns true iff value == expect1 ? 1 : 0
template <class M> HD void mach_canon_be(uint8_t* be, const uint32_t* limbs, bool* is_zero, bool* is_one) {
    constexpr int NW = M::FP_BYTES / 4, L = M::L, W = M::W;
    constexpr uint32_t MASK = (1u << W) - 1;
    // conditional subtraction of p on W-bit limbs
    uint32_t d[L];
    int borrow = 0;
#pragma unroll
    for (int i = 0; i < L; i++) {
        int x = (int)limbs[i] - (int)M::p(i) - borrow;
        borrow = x < 0;
        d[i] = (uint32_t)x & MASK;
    }
    uint32_t v[L];
#pragma unroll
    for (int i = 0; i < L; i++) v[i] = borrow ? limbs[i] : d[i];
    // repack to 32-bit words
    uint32_t w[NW];
    uint32_t nz = 0;
#pragma unroll
    for (int k = 0; k < NW; k++) {
        const int bit = 32 * k, li = bit / W, sh = bit % W;
        unsigned long long x = (unsigned long long)v[li] >> sh;
        if (li + 1 < L) x |= (unsigned long long)v[li + 1] << (W - sh);
        if (li + 2 < L && 2 * W - sh < 32) x |= (unsigned long long)v[li + 2] << (2 * W - sh);
        w[k] = (uint32_t)x;
        nz |= (k == 0) ? (w[k] ^ 1u) : w[k];
    }
    uint32_t any = 0;
#pragma unroll
    for (int k = 0; k < NW; k++) {
        any |= w[k];
        uint8_t* q = be + 4 * (NW - 1 - k);
        q[0] = (uint8_t)(w[k] >> 24);
        q[1] = (uint8_t)(w[k] >> 16);
        q[2] = (uint8_t)(w[k] >> 8);
        q[3] = (uint8_t)w[k];
    }
    *is_zero = any == 0;
    *is_one = nz == 0;
}

// ---- 32-bit word helpers for the single-lane inverse
template <class M> HD void mach_limbs_to_words(uint32_t* w, const uint32_t* v) {
    constexpr int NW = M::FP_BYTES / 4, L = M::L, W = M::W;
#pragma unroll
    for (int k = 0; k < NW; k++) {
        const int bit = 32 * k, li = bit / W, sh = bit % W;
        unsigned long long x = (unsigned long long)v[li] >> sh;
        if (li + 1 < L) x |= (unsigned long long)v[li + 1] << (W - sh);
        if (li + 2 < L && 2 * W - sh < 32) x |= (unsigned long long)v[li + 2] << (2 * W - sh);
        w[k] = (uint32_t)x;
    }
}
template <class M> HD void mach_words_to_limbs(uint32_t* out, const uint32_t* win) {
    constexpr int NW = M::FP_BYTES / 4, L = M::L, W = M::W;
    constexpr uint32_t MASK = (1u << W) - 1;
    uint32_t w[NW + 1];
#pragma unroll
    for (int i = 0; i < NW; i++) w[i] = win[i];
    w[NW] = 0;
#pragma unroll
    for (int i = 0; i < L; i++) {
        const int bit = W * i, wi = bit >> 5, sh = bit & 31;
        const uint32_t lo = w[wi], hi = w[wi + 1];
        out[i] = (sh ? ((lo >> sh) | (hi << (32 - sh))) : lo) & MASK;
    }
}
template <int NW> HD uint32_t w_add(uint32_t* r, const uint32_t* a, const uint32_t* b) {
    unsigned long long c = 0;
#pragma unroll
    for (int i = 0; i < NW; i++) { c += (unsigned long long)a[i] + b[i]; r[i] = (uint32_t)c; c >>= 32; }
    return (uint32_t)c;
}
template <int NW> HD uint32_t w_sub(uint32_t* r, const uint32_t* a, const uint32_t* b) {
    long long c = 0;
#pragma unroll
    for (int i = 0; i < NW; i++) { c += (long long)a[i] - (long long)b[i]; r[i] = (uint32_t)c; c >>= 32; }
    return c != 0;  // borrow
}
template <int NW> HD void w_shr1(uint32_t* a, uint32_t top) {
#pragma unroll
    for (int i = 0; i < NW - 1; i++) a[i] = (a[i] >> 1) | (a[i + 1] << 31);
    a[NW - 1] = (a[NW - 1] >> 1) | (top << 31);
}
template <int NW> HD bool w_geq(const uint32_t* a, const uint32_t* b) {
    for (int i = NW - 1; i >= 0; i--) {
        if (a[i] != b[i]) return a[i] > b[i];
    }
    return true;
}
template <int NW> HD bool w_is_one(const uint32_t* a) {
    uint32_t x = a[0] ^ 1u;
#pragma unroll
    for (int i = 1; i < NW; i++) x |= a[i];
    return x == 0;
}
// out = in^-1 mod p; in: plain-domain value < 2p as L limbs, out: L limbs < p (0 -> 0).
// Binary extended Euclid (right-shift variant, p odd): ~2 log2(p) shift/subtract steps on one lane instead of
// the ~380 dependent Montgomery multiplications of a Fermat inversion.
template <class M> HD void mach_inv(uint32_t* out, const uint32_t* in) {
    constexpr int NW = M::FP_BYTES / 4, L = M::L;
    uint32_t P[NW], u[NW], v[NW], x1[NW], x2[NW], t[NW];
    uint32_t pl[L];
#pragma unroll
    for (int i = 0; i < L; i++) pl[i] = M::p(i);
    mach_limbs_to_words<M>(P, pl);
    mach_limbs_to_words<M>(u, in);
    if (!w_sub<NW>(t, u, P)) {
#pragma unroll
        for (int i = 0; i < NW; i++) u[i] = t[i];
    }
    uint32_t nz = 0;
#pragma unroll
    for (int i = 0; i < NW; i++) { v[i] = P[i]; x1[i] = 0; x2[i] = 0; nz |= u[i]; }
    x1[0] = 1;
    if (nz != 0) {
        for (int guard = 0; guard < 4 * 32 * NW && !w_is_one<NW>(u) && !w_is_one<NW>(v); guard++) {
            if ((u[0] & 1u) == 0) {
                w_shr1<NW>(u, 0);
                uint32_t c = 0;
                if (x1[0] & 1u) c = w_add<NW>(x1, x1, P);
                w_shr1<NW>(x1, c);
            } else if ((v[0] & 1u) == 0) {
                w_shr1<NW>(v, 0);
                uint32_t c = 0;
                if (x2[0] & 1u) c = w_add<NW>(x2, x2, P);
                w_shr1<NW>(x2, c);
            } else if (w_geq<NW>(u, v)) {
                w_sub<NW>(u, u, v);
                if (w_sub<NW>(x1, x1, x2)) w_add<NW>(x1, x1, P);
            } else {
                w_sub<NW>(v, v, u);
                if (w_sub<NW>(x2, x2, x1)) w_add<NW>(x2, x2, P);
            }
        }
        const bool use1 = w_is_one<NW>(u);
#pragma unroll
        for (int i = 0; i < NW; i++) t[i] = use1 ? x1[i] : x2[i];
    } else {
#pragma unroll
        for (int i = 0; i < NW; i++) t[i] = 0;
    }
    mach_words_to_limbs<M>(out, t);
}

// ---- tables on the device: one struct of pointers per slot file
struct MachTables {
    const uint32_t* consts;   // [NCONST][L]
    const uint32_t* hdr;      // [NPHASE]
    const uint16_t* rec;      // [NPHASE][LANES][REC]
};

// run one phase for one lane (device) -- the caller synchronises the warp afterwards
template <class M> HD void mach_phase_lane(const MachView<M>& mv, const MachTables& tb, uint32_t phase, int lane) {
    const uint32_t h = tb.hdr[phase];
    const int T = (h >> 8) & 0xFF;
    const uint16_t* rec = tb.rec + ((size_t)phase * M::LANES + lane) * M::REC;
    uint32_t out[M::L];
    const uint32_t kind = h & 0xFF;
    if (kind == 2) {  // single-lane modular inverse
        if (rec[0] != MIDLE) {
            uint32_t in[M::L];
            mach_load<M>(in, mv, rec[1]);
            mach_inv<M>(out, in);
            mach_store<M>(mv, rec[0], out);
        }
        return;
    }
#ifdef __CUDA_ARCH__
    if constexpr (M::SIGNED && M::LANES == 32) {
        if (kind == 3) {   // the whole warp is here: the phase kind is uniform
            unsigned long long acc[2 * M::L];
            mach_kdot_acc<M>(mv, rec, T, acc);
            mach_kdot_combine<M>(acc, (h >> 16) & 0xFF, lane);
            mach_kdot_finish<M>(out, acc, h >> 24);
            if (rec[0] != MIDLE) mach_store<M>(mv, rec[0], out);
            return;
        }
    }
#endif
    if (kind == 0) mach_dot<M>(mv, rec, T, h >> 24, out);
    else mach_lin<M>(mv, rec, T, out);
    if (rec[0] != MIDLE) mach_store<M>(mv, rec[0], out);
}

}  // namespace bgls
