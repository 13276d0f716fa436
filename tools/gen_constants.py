"""Generates bgls_b200/csrc/curve_params.cuh: Montgomery-form constants (32-bit limbs) for the
CUDA kernels.  Self-contained integer arithmetic (does NOT import oracle/): the product's
constants are derived independently of the checker.   python tools/gen_constants.py
"""
import os


def limbs(v, n):
    return [(v >> (32 * i)) & 0xFFFFFFFF for i in range(n)]


def arr(vals):
    return "{" + ", ".join("0x%08xu" % x for x in vals) + "}"


def f2mul(a, b, p):
    return ((a[0] * b[0] - a[1] * b[1]) % p, (a[0] * b[1] + a[1] * b[0]) % p)


def f2pow(a, e, p):
    r = (1, 0)
    while e:
        if e & 1:
            r = f2mul(r, a, p)
        a = f2mul(a, a, p)
        e >>= 1
    return r


def f2inv(a, p):
    n = pow(a[0] * a[0] + a[1] * a[1], -1, p)
    return (a[0] * n % p, -a[1] * n % p)


def emit(name, p, r, xi, b, twist, n, extra, extra_consts=()):
    R = 1 << (32 * n)
    mont = lambda v: limbs(v * R % p, n)
    mont2 = lambda a: mont(a[0]) + mont(a[1])
    n0 = (-pow(p, -1, 1 << 32)) % (1 << 32)
    if twist == "D":
        b2 = f2mul((b, 0), f2inv(xi, p), p)
    else:
        b2 = f2mul((b, 0), xi, p)
    b2x3 = (3 * b2[0] % p, 3 * b2[1] % p)
    g1 = f2pow(xi, (p - 1) // 6, p)
    gam = [(1, 0)]
    for k in range(1, 6):
        gam.append(f2mul(gam[-1], g1, p))
    # p^2-Frobenius: coefficient k scaled by N(gamma_1)^k = gamma_1^(k(p+1)) in Fp
    g2 = [f2pow(gam[k], p + 1, p) for k in range(6)]
    assert all(x[1] == 0 for x in g2)
    # p^3-Frobenius: conj then gamma_1^(k (p^2+p+1))
    g3 = [f2pow(gam[k], p * p + p + 1, p) for k in range(6)]
    out = []
    U = name.upper()
    out.append(f"// ---- {name}: p = 0x{p:x}")
    out.append(f"DEVCONST uint32_t {U}_P[{n}] = {arr(limbs(p, n))};")
    out.append(f"DEVCONST uint32_t {U}_R1[{n}] = {arr(mont(1))};")
    out.append(f"DEVCONST uint32_t {U}_R2[{n}] = {arr(limbs(R * R % p, n))};")
    out.append(f"DEVCONST uint32_t {U}_HALF[{n}] = {arr(mont(pow(2, -1, p)))};")
    out.append(f"DEVCONST uint32_t {U}_B1[{n}] = {arr(mont(b))};")
    out.append(f"DEVCONST uint32_t {U}_B1X3[{n}] = {arr(mont(3 * b))};")
    out.append(f"DEVCONST uint32_t {U}_B2[{2 * n}] = {arr(mont2(b2))};")
    out.append(f"DEVCONST uint32_t {U}_B2X3[{2 * n}] = {arr(mont2(b2x3))};")
    out.append(f"DEVCONST uint32_t {U}_GAMMA1[6][{2 * n}] = {{" + ", ".join(arr(mont2(g)) for g in gam) + "};")
    out.append(f"DEVCONST uint32_t {U}_GAMMA2[6][{n}] = {{" + ", ".join(arr(mont(g[0])) for g in g2) + "};")
    out.append(f"DEVCONST uint32_t {U}_GAMMA3[6][{2 * n}] = {{" + ", ".join(arr(mont2(g)) for g in g3) + "};")
    out.append(f"DEVCONST uint32_t {U}_PM2[{n}] = {arr(limbs(p - 2, n))};  // exponent for Fermat inversion")
    out.append(f"DEVCONST uint32_t {U}_PP1D4[{n}] = {arr(limbs((p + 1) // 4, n))};  // square root exponent (p = 3 mod 4)")
    out.append(f"DEVCONST uint32_t {U}_PM1D2[{n}] = {arr(limbs((p - 1) // 2, n))};  // Euler criterion exponent, also the parity threshold")
    for cname, cval in extra_consts:
        out.append(f"DEVCONST uint32_t {U}_{cname}[{n}] = {arr(mont(cval))};")
    out.append(f"struct {name} {{")
    out.append(f"    static constexpr int N = {n};")
    out.append(f"    static constexpr bool IS_BN = {'true' if twist == 'D' else 'false'};")
    out.append(f"    static constexpr uint32_t N0 = 0x{n0:08x}u;")
    out.append("    HD static constexpr uint32_t p(int i) {")
    out.append(f"        constexpr uint32_t v[{n}] = {arr(limbs(p, n))};")
    out.append("        return v[i];")
    out.append("    }")
    for t in ["P", "R1", "R2", "HALF", "B1", "B1X3", "B2", "B2X3", "PM2", "PP1D4", "PM1D2"] + [c for c, _ in extra_consts]:
        out.append(f"    HD static const uint32_t* {t}() {{ return {U}_{t}; }}")
    for t in ("GAMMA1", "GAMMA2", "GAMMA3"):
        out.append(f"    HD static const uint32_t* {t}(int k) {{ return {U}_{t}[k]; }}")
    out.extend(extra)
    out.append("};")
    return "\n".join(out)


def main():
    bn_p = 21888242871839275222246405745257275088696311157297823662689037894645226208583
    bn_r = 21888242871839275222246405745257275088548364400416034343698204186575808495617
    u = 4965661367192848881
    bl_p = 0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab
    bl_r = 52435875175126190479447740508185965837690552500527637822603658699938581184513
    x = 0xD201000000010000
    s = 6 * u + 2
    assert s.bit_length() == 65
    c = (x + 1) ** 2 // 3
    assert 3 * c == (x + 1) ** 2
    bn_extra = [
        f"    static constexpr unsigned long long LOOP_LO = 0x{s & (2**64 - 1):x}ull;  // 6u+2 without its leading bit (bit 64)",
        "    static constexpr int LOOP_TOP = 64;  // number of iterations after the leading bit",
        f"    static constexpr unsigned long long U = 0x{u:x}ull;  // BN parameter u",
        "    static constexpr unsigned long long C_LO = 0, C_HI = 0;  // unused on BN",
        "    static constexpr int FP_BYTES = 32;",
    ]
    bl_extra = [
        f"    static constexpr unsigned long long LOOP_LO = 0x{x:x}ull;  // |x|",
        "    static constexpr int LOOP_TOP = 63;",
        f"    static constexpr unsigned long long U = 0x{x:x}ull;  // |x|",
        f"    static constexpr unsigned long long C_LO = 0x{c & (2**64 - 1):x}ull, C_HI = 0x{c >> 64:x}ull;  // (|x|+1)^2/3",
        "    static constexpr int FP_BYTES = 48;",
    ]
    # Fouque-Tibouchi / SvdW hash constants and the G1 generator (curves/bls12_381.go:333-346)
    bl_consts = [
        ("FT_SQRT_NEG3", 1586958781458431025242759403266842894121773480562120986020912974854563298150952611241517463240701),
        ("FT_Z", 793479390729215512621379701633421447060886740281060493010456487427281649075476305620758731620350),
        ("FT_ROOT1", 248294325734266649657405162895821171812231848760181225578082735178502750823719347628762635478508544819911854747095),
        ("FT_ROOT2", 3754115229487400743760384662840082984744650971178826659753975400945528899667118516813924993650507119217982417812692),
        ("G1X", 0x17f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac586c55e83ff97a1aeffb3af00adb22c6bb),
        ("G1Y", 0x08b3f481e3aaa0f1a09e30ed741d8ae4fcf5e095d5d00af600db18cb2c04b3edd03cc744a2888ae40caa232946c5e7e1),
        ("TWO376", 1 << 376),
    ]
    hdr = [
        "// GENERATED by tools/gen_constants.py -- do not edit.",
        "// Montgomery constants, R = 2^(32 N); 32-bit little-endian limbs.",
        "#pragma once",
        "#include <cstdint>",
        '#include "arith.cuh"',
        "namespace bgls {",
        emit("BN254", bn_p, bn_r, (9, 1), 3, "D", 8, bn_extra, [("TWO248", 1 << 248)]),
        emit("BLS381", bl_p, bl_r, (1, 1), 4, "M", 12, bl_extra, bl_consts),
        "}  // namespace bgls",
        "",
    ]
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bgls_b200", "csrc", "curve_params.cuh")
    with open(path, "w") as f:
        f.write("\n".join(hdr))
    print("wrote", path)


if __name__ == "__main__":
    main()
