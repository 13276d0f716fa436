// Carry-chain primitives for multi-limb modular arithmetic on sm_100a.
//
// Device build (nvcc): each primitive is one PTX instruction (or one lo/hi pair that ptxas fuses
// into a single IMAD.WIDE.U32[.X] with predicate carry -- verified with cuobjdump -sass).
// Host build (g++, tests/host_emul only): the same primitives are emulated with an explicit
// carry flag so every algorithm above this layer can be unit-tested on a CPU-only machine
// against the oracle.  The host build is test scaffolding, never linked into the product.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define HD __device__ __forceinline__
#define HDNI __device__ __noinline__
#define DEVCONST static __device__ __constant__ const
#else
#define HD inline
#define HDNI inline
#define DEVCONST static const
#endif

namespace bgls {

#if defined(__CUDACC__)

HD void mul_wide(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) {
    asm volatile("mul.lo.u32 %0, %2, %3; mul.hi.u32 %1, %2, %3;" : "=r"(lo), "=r"(hi) : "r"(a), "r"(b));
}
// (hi:lo) += a*b, carry out
HD void mad_wide_cc(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) {
    asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(lo), "+r"(hi) : "r"(a), "r"(b));
}
// (hi:lo) += a*b + carry, carry out
HD void madc_wide_cc(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) {
    asm volatile("madc.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(lo), "+r"(hi) : "r"(a), "r"(b));
}
// (hi:lo) = a*b + (chi:clo) + carry, carry out
HD void madc_wide_cc3(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b, uint32_t clo, uint32_t chi) {
    asm volatile("madc.lo.cc.u32 %0, %2, %3, %4; madc.hi.cc.u32 %1, %2, %3, %5;"
                 : "=r"(lo), "=r"(hi) : "r"(a), "r"(b), "r"(clo), "r"(chi));
}
HD void add_cc(uint32_t& d, uint32_t a, uint32_t b) { asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); }
HD void addc_cc(uint32_t& d, uint32_t a, uint32_t b) { asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); }
HD void addc(uint32_t& d, uint32_t a, uint32_t b) { asm volatile("addc.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); }
HD void sub_cc(uint32_t& d, uint32_t a, uint32_t b) { asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); }
HD void subc_cc(uint32_t& d, uint32_t a, uint32_t b) { asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); }
HD void subc(uint32_t& d, uint32_t a, uint32_t b) { asm volatile("subc.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); }

#else  // ---- host emulation (tests only)

static thread_local uint32_t g_cf = 0;

inline void mul_wide(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) {
    uint64_t x = (uint64_t)a * b;
    lo = (uint32_t)x;
    hi = (uint32_t)(x >> 32);
}
inline void mad_wide_cc(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) {
    unsigned __int128 x = (unsigned __int128)((uint64_t)a * b) + (((uint64_t)hi << 32) | lo);
    lo = (uint32_t)x;
    hi = (uint32_t)(x >> 32);
    g_cf = (uint32_t)(x >> 64);
}
inline void madc_wide_cc(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) {
    unsigned __int128 x = (unsigned __int128)((uint64_t)a * b) + (((uint64_t)hi << 32) | lo) + g_cf;
    lo = (uint32_t)x;
    hi = (uint32_t)(x >> 32);
    g_cf = (uint32_t)(x >> 64);
}
inline void madc_wide_cc3(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b, uint32_t clo, uint32_t chi) {
    unsigned __int128 x = (unsigned __int128)((uint64_t)a * b) + (((uint64_t)chi << 32) | clo) + g_cf;
    lo = (uint32_t)x;
    hi = (uint32_t)(x >> 32);
    g_cf = (uint32_t)(x >> 64);
}
inline void add_cc(uint32_t& d, uint32_t a, uint32_t b) { uint64_t x = (uint64_t)a + b; d = (uint32_t)x; g_cf = (uint32_t)(x >> 32); }
inline void addc_cc(uint32_t& d, uint32_t a, uint32_t b) { uint64_t x = (uint64_t)a + b + g_cf; d = (uint32_t)x; g_cf = (uint32_t)(x >> 32); }
inline void addc(uint32_t& d, uint32_t a, uint32_t b) { d = a + b + g_cf; }
inline void sub_cc(uint32_t& d, uint32_t a, uint32_t b) { uint64_t x = (uint64_t)a - b; d = (uint32_t)x; g_cf = (uint32_t)(x >> 32) & 1; }
inline void subc_cc(uint32_t& d, uint32_t a, uint32_t b) { uint64_t x = (uint64_t)a - b - g_cf; d = (uint32_t)x; g_cf = (uint32_t)(x >> 32) & 1; }
inline void subc(uint32_t& d, uint32_t a, uint32_t b) { d = a - b - g_cf; }

#endif

}  // namespace bgls
