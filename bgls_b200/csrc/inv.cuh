// Modular inversion by a binary extended GCD (Kaliski's almost inverse with the halvings of a run of zero bits
// merged into one multi-word shift): ~0.7 log2(p) iterations of one subtraction, one addition and two shifts on N
// limbs -- about a tenth of the instructions of the Fermat power a^(p-2) it replaces (570 Montgomery
// multiplications for a 381-bit prime).  Thread serial; used by the one inversion at the end of AggregatePoints
// (agg.cuh) and by fp_inv (field.cuh: affine conversions, hash-to-G1, the codecs, the thread engine's final
// exponentiation).
#pragma once
#include <cstdint>

#include "arith.cuh"

namespace bgls {

#if defined(__CUDACC__)
#define INV_NOINLINE __device__ __noinline__
#else
#define INV_NOINLINE inline
#endif

template <int N> HD bool mpw_is_zero(const uint32_t* a) {
    uint32_t any = 0;
#pragma unroll
    for (int i = 0; i < N; i++) any |= a[i];
    return any == 0;
}
template <int N> HD uint32_t mpw_sub(uint32_t* r, const uint32_t* a, const uint32_t* b) {   // returns the borrow mask
    uint32_t br;
    sub_cc(r[0], a[0], b[0]);
#pragma unroll
    for (int i = 1; i < N; i++) subc_cc(r[i], a[i], b[i]);
    subc(br, 0, 0);
    return br;
}
template <int N> HD void mpw_add(uint32_t* r, const uint32_t* a, const uint32_t* b) {
    add_cc(r[0], a[0], b[0]);
#pragma unroll
    for (int i = 1; i < N - 1; i++) addc_cc(r[i], a[i], b[i]);
    addc(r[N - 1], a[N - 1], b[N - 1]);
}
HD int inv_ctz31(uint32_t x) {   // trailing zeros of x, at most 31 (x = 0 -> 31)
#if defined(__CUDA_ARCH__)
    return __ffs((int)(x | 0x80000000u)) - 1;
#else
    return __builtin_ctz(x | 0x80000000u);
#endif
}
template <int N> HD void mpw_shr(uint32_t* a, int t) {   // 1 <= t <= 31
#pragma unroll
    for (int i = 0; i < N - 1; i++) {
#if defined(__CUDA_ARCH__)
        a[i] = __funnelshift_r(a[i], a[i + 1], t);
#else
        a[i] = (a[i] >> t) | (a[i + 1] << (32 - t));
#endif
    }
    a[N - 1] >>= t;
}
template <int N> HD void mpw_shl(uint32_t* a, int t) {   // 1 <= t <= 31
#pragma unroll
    for (int i = N - 1; i > 0; i--) {
#if defined(__CUDA_ARCH__)
        a[i] = __funnelshift_l(a[i - 1], a[i], t);
#else
        a[i] = (a[i] << t) | (a[i - 1] >> (32 - t));
#endif
    }
    a[0] <<= t;
}

// x = a^-1 2^k mod p, 0 < a < p (N plain limbs), returns k in [log2 p, 2 log2 p]
template <class C> INV_NOINLINE int mp_almost_inv(uint32_t* x, const uint32_t* a) {
    constexpr int N = C::N;
    uint32_t u[N], v[N], r[N], s[N], d[N];
#pragma unroll
    for (int i = 0; i < N; i++) { u[i] = C::p(i); v[i] = a[i]; r[i] = 0; s[i] = i == 0 ? 1u : 0u; }
    int k = 0;
    // invariant of the loop: u and v odd.  (r stays 0 while v is made odd.)
    while (!(v[0] & 1u)) {
        if (mpw_is_zero<N>(v)) break;   // a = 0: not invertible, the result is meaningless
        const int t = inv_ctz31(v[0]);
        mpw_shr<N>(v, t);
        k += t;
    }
#pragma unroll 1
    for (;;) {
        const uint32_t br = mpw_sub<N>(d, v, u);          // v - u (even)
        if (!br) {                                        // v >= u:  v <- (v - u) / 2^t,  s <- s + r,  r <- r 2^t
            if (d[0] == 0 && mpw_is_zero<N>(d)) break;    // v = u = gcd
            mpw_add<N>(s, s, r);
            do {
                const int t = inv_ctz31(d[0]);
                mpw_shr<N>(d, t);
                mpw_shl<N>(r, t);
                k += t;
            } while (!(d[0] & 1u));
#pragma unroll
            for (int i = 0; i < N; i++) v[i] = d[i];
        } else {                                          // u > v:  u <- (u - v) / 2^t,  r <- r + s,  s <- s 2^t
            mpw_sub<N>(u, u, v);
            mpw_add<N>(r, r, s);
            do {
                const int t = inv_ctz31(u[0]);
                mpw_shr<N>(u, t);
                mpw_shl<N>(s, t);
                k += t;
            } while (!(u[0] & 1u));
        }
    }
    // the step that takes v to zero: s <- s + r, r <- 2 r
    mpw_shl<N>(r, 1);
    k += 1;
    // r = -a^-1 2^k mod p, in [0, 2p)
    uint32_t pp[N];
#pragma unroll
    for (int i = 0; i < N; i++) pp[i] = C::p(i);
    if (!mpw_sub<N>(d, r, pp)) {
#pragma unroll
        for (int i = 0; i < N; i++) r[i] = d[i];
    }
    mpw_sub<N>(x, pp, r);
    return k;
}
// Jacobi symbol (a / p) of 0 <= a < p (N plain limbs; for a Montgomery residue a R the symbol is that of a, R being a
// square): -1, 0 (a = 0) or 1.  Binary algorithm -- subtract the smaller of two odd numbers from the larger, strip the
// zero bits with the (2 / n) rule, quadratic reciprocity whenever numerator and denominator change roles -- about 380
// iterations of one subtraction and one shift for a 381-bit prime: the residuosity test of hash-to-G1
// (/root/reference/curves/hash.go:254-265, a ~476-multiplication exponentiation there) for the price of ~27.
template <class C> INV_NOINLINE int mp_jacobi(const uint32_t* a) {
    constexpr int N = C::N;
    uint32_t A[N], B[N], d[N];
#pragma unroll
    for (int i = 0; i < N; i++) { A[i] = a[i]; B[i] = C::p(i); }
    if (mpw_is_zero<N>(A)) return 0;
    int t = 1;
    bool num_is_a = true;                         // which array is the numerator of the current symbol
    while (!(A[0] & 1u)) {                        // (2 / p)^tz
        const int tz = inv_ctz31(A[0]);
        mpw_shr<N>(A, tz);
        if ((tz & 1) && (((B[0] & 7u) == 3u) || ((B[0] & 7u) == 5u))) t = -t;
    }
#pragma unroll 1
    for (;;) {
        const uint32_t br = mpw_sub<N>(d, A, B);  // A - B
        if (!br) {
            if (d[0] == 0 && mpw_is_zero<N>(d)) break;                     // A = B = gcd
            if (!num_is_a) {                                               // (B / A) -> (A / B): reciprocity
                if ((A[0] & 3u) == 3u && (B[0] & 3u) == 3u) t = -t;
                num_is_a = true;
            }
            do {
                const int tz = inv_ctz31(d[0]);
                mpw_shr<N>(d, tz);
                if ((tz & 1) && (((B[0] & 7u) == 3u) || ((B[0] & 7u) == 5u))) t = -t;
            } while (!(d[0] & 1u));
#pragma unroll
            for (int i = 0; i < N; i++) A[i] = d[i];
        } else {
            if (num_is_a) {
                if ((A[0] & 3u) == 3u && (B[0] & 3u) == 3u) t = -t;
                num_is_a = false;
            }
            mpw_sub<N>(B, B, A);
            do {
                const int tz = inv_ctz31(B[0]);
                mpw_shr<N>(B, tz);
                if ((tz & 1) && (((A[0] & 7u) == 3u) || ((A[0] & 7u) == 5u))) t = -t;
            } while (!(B[0] & 1u));
        }
    }
    return (A[0] == 1u && mpw_is_zero<N - 1>(A + 1)) ? t : 0;
}

// the two exponents of x 2^-k = mont(mont(x, 2^j1), 2^j2), mont(a, b) = a b 2^(-32 N):  j1 + j2 = 64 N - k
template <int N> HD void inv_shift_limbs(uint32_t* e1, uint32_t* e2, int k) {
    const int j = 64 * N - k, j1 = j < 32 * N - 1 ? j : 32 * N - 1, j2 = j - j1;
#pragma unroll
    for (int i = 0; i < N; i++) {
        e1[i] = (j1 >> 5) == i ? 1u << (j1 & 31) : 0u;
        e2[i] = (j2 >> 5) == i ? 1u << (j2 & 31) : 0u;
    }
}

}  // namespace bgls
