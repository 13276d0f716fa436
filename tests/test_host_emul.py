"""The device arithmetic templates (bgls_b200/csrc/*.cuh) compiled for the host with emulated carry
primitives (tests/host_emul/emul.cpp, test scaffolding) against the oracle: every algorithm the GPU
runs -- saturated tower arithmetic, group law, and the dot-product machine interpreter with its
generated programs -- is proven on a CPU-only machine."""
import ctypes
import os
import random
import subprocess

import pytest

from oracle import c_oracle as C
from parity_util import CURVES, GOLD

HERE = os.path.dirname(os.path.abspath(__file__))
EMU = os.path.join(HERE, "host_emul")


@pytest.fixture(scope="module")
def E():
    so = os.path.join(EMU, "libemul.so")
    src = os.path.join(EMU, "emul.cpp")
    csrc = os.path.join(os.path.dirname(HERE), "bgls_b200", "csrc")
    newest = max(os.path.getmtime(os.path.join(csrc, f)) for f in os.listdir(csrc))
    if not os.path.exists(so) or os.path.getmtime(so) < max(newest, os.path.getmtime(src)):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", so, src])
    L = ctypes.CDLL(so)
    u8p = ctypes.c_char_p
    L.emu_pairing_product.argtypes = [ctypes.c_int, u8p, u8p, ctypes.c_size_t, u8p, ctypes.c_int]
    L.emu_aggregate.argtypes = [ctypes.c_int, ctypes.c_int, u8p, ctypes.c_size_t, u8p]
    L.emu_scale.argtypes = [ctypes.c_int, ctypes.c_int, u8p, u8p, ctypes.c_size_t, u8p]
    L.emu_fpops.argtypes = [ctypes.c_int, u8p, u8p, u8p]
    L.emu_mach_pairing_product.argtypes = [ctypes.c_int, u8p, u8p, ctypes.c_size_t, u8p, ctypes.c_int, ctypes.c_void_p]
    L.emu_mach_pairing_product_p.argtypes = [ctypes.c_int, u8p, u8p, ctypes.c_size_t, u8p, ctypes.c_int]
    return L


@pytest.mark.parametrize("cid,c", CURVES)
def test_fp_ops_even_odd_montgomery(E, cid, c):
    rng = random.Random(3)
    nb = c.nbytes
    cases = [(c.p - 1, c.p - 1), (0, 5), (1, 1)] + [(rng.randrange(c.p), rng.randrange(c.p)) for _ in range(200)]
    for a, b in cases:
        out = ctypes.create_string_buffer(4 * nb)
        E.emu_fpops(cid, a.to_bytes(nb, "big"), b.to_bytes(nb, "big"), out)
        o = out.raw
        assert int.from_bytes(o[:nb], "big") == a * b % c.p
        assert int.from_bytes(o[nb:2 * nb], "big") == (a + b) % c.p
        assert int.from_bytes(o[2 * nb:3 * nb], "big") == (a - b) % c.p
        if a:
            assert int.from_bytes(o[3 * nb:], "big") == pow(a, -1, c.p)


@pytest.mark.parametrize("cid,c", CURVES)
def test_thread_engine_templates(E, cid, c):
    g = GOLD[c.name]
    g1, g2, n = bytes.fromhex(g["g1"]), bytes.fromhex(g["g2"]), g["n"]
    nb = c.nbytes
    out = ctypes.create_string_buffer(12 * nb)
    E.emu_pairing_product(cid, g1, g2, n, out, 1)
    assert out.raw.hex() == g["product_gt"]
    E.emu_pairing_product(cid, g1, g2, n, out, 0)
    assert out.raw == C.miller_product(cid, g1, g2, n)  # same formulas as the C oracle: raw values agree too
    # shared Miller accumulator (K = 4 pairs per thread, ragged tail, an infinity pair in the middle)
    E.emu_pairing_product_shared.argtypes = E.emu_pairing_product.argtypes
    E.emu_pairing_product_shared(cid, g1, g2, n, out, 1)
    assert out.raw.hex() == g["product_gt"]
    G1 = g1[:2 * nb] + c.marshal_g1(None) + g1[2 * nb:]
    G2 = g2[:4 * nb] + c.marshal_g2(c.g2) + g2[4 * nb:]
    E.emu_pairing_product_shared(cid, G1, G2, n + 1, out, 1)
    assert out.raw.hex() == g["product_gt"]
    E.emu_pairing_product_shared(cid, g1, g2, n, out, 0)
    assert C.fp12_product(cid, out.raw, 1, True).hex() == g["product_gt"]
    rng = random.Random(8)
    for grp, key in ((1, "g1"), (2, "g2")):
        blob = bytes.fromhex(g[key])
        rec = 2 * grp * nb
        o = ctypes.create_string_buffer(rec)
        E.emu_aggregate(cid, grp, blob, n, o)
        assert o.raw.hex() == g["sum_" + key]
        sc = b"".join(rng.randrange(c.r).to_bytes(32, "big") for _ in range(n))
        o = ctypes.create_string_buffer(rec * n)
        E.emu_scale(cid, grp, blob, sc, n, o)
        assert o.raw == C.scale_points(cid, grp, blob, sc, n)


@pytest.mark.parametrize("cid,c", CURVES)
def test_machine_interpreter_and_programs(E, cid, c):
    g = GOLD[c.name]
    g1, g2, n = bytes.fromhex(g["g1"]), bytes.fromhex(g["g2"]), g["n"]
    nb = c.nbytes
    out = ctypes.create_string_buffer(12 * nb)
    assert E.emu_mach_pairing_product(cid, g1, g2, n, out, 1, None) == 0
    assert out.raw.hex() == g["product_gt"]
    # raw Miller product (EXPORT program): equal to the oracle's after the oracle's final exponentiation
    E.emu_mach_pairing_product(cid, g1, g2, n, out, 0, None)
    assert C.fp12_product(cid, out.raw, 1, True).hex() == g["product_gt"]
    # valid aggregate signature + an infinity pair -> identity
    a = g["agg"]
    hs = [bytes.fromhex(h) for h in a["hashes"]]
    pks = [bytes.fromhex(h) for h in a["pubkeys"]]
    neg = c.marshal_g1(c.g1_neg(c.unmarshal_g1(bytes.fromhex(a["sig"]))))
    G1 = b"".join(hs) + neg + c.marshal_g1(None)
    G2 = b"".join(pks) + bytes.fromhex(a["g2gen"]) + c.marshal_g2(c.g2)
    assert E.emu_mach_pairing_product(cid, G1, G2, 5, out, 1, None) == 1
    assert out.raw == c.marshal_gt(c.fp12_one)
    assert E.emu_mach_pairing_product(cid, b"", b"", 0, out, 1, None) == 1


@pytest.mark.parametrize("cid,c", CURVES)
def test_pipelined_miller_program(E, cid, c):
    """The 32-lane pipelined Miller program (signed DOT terms, slot file P) through the real interpreter code."""
    g = GOLD[c.name]
    g1, g2, n = bytes.fromhex(g["g1"]), bytes.fromhex(g["g2"]), g["n"]
    nb = c.nbytes
    out = ctypes.create_string_buffer(12 * nb)
    assert E.emu_mach_pairing_product_p(cid, g1, g2, n, out, 1) == 0
    assert out.raw.hex() == g["product_gt"]
    a = g["agg"]
    hs = [bytes.fromhex(h) for h in a["hashes"]]
    pks = [bytes.fromhex(h) for h in a["pubkeys"]]
    neg = c.marshal_g1(c.g1_neg(c.unmarshal_g1(bytes.fromhex(a["sig"]))))
    G1 = b"".join(hs) + neg + c.marshal_g1(None)
    G2 = b"".join(pks) + bytes.fromhex(a["g2gen"]) + c.marshal_g2(c.g2)
    assert E.emu_mach_pairing_product_p(cid, G1, G2, 5, out, 1) == 1
    assert out.raw == c.marshal_gt(c.fp12_one)
    rng = random.Random(23)
    for _ in range(3):   # random pairs: the raw Miller value differs from the oracle's by subfield factors only
        P = c.g1_mul(c.g1, rng.randrange(c.r))
        Q = c.g2_mul(c.g2, rng.randrange(c.r))
        assert E.emu_mach_pairing_product_p(cid, c.marshal_g1(P), c.marshal_g2(Q), 1, out, 1) == 0
        assert out.raw == c.marshal_gt(c.pair(P, Q))


@pytest.mark.parametrize("cid,c", CURVES)
def test_hash_to_g1_reference_vectors(E, cid, c):
    """The device hash-to-G1 functions against the reference's own vectors (curves/curve_test.go:210-244,
    curves/altbn128_test.go:13-24, curves/bls12_test.go:57-67) and against the oracle on ragged lengths."""
    import base64
    E.emu_hash_to_g1.argtypes = [ctypes.c_int, ctypes.c_char_p, ctypes.c_size_t, ctypes.c_char_p]
    nb = c.nbytes

    def H(m):
        out = ctypes.create_string_buffer(2 * nb)
        E.emu_hash_to_g1(cid, m, len(m), out)
        return out.raw
    fname = "altbn128G1Hash.dat" if cid == 0 else "bls12G1Hash.dat"
    for line in open(os.path.join(HERE, "golden", fname)):
        m, pt = line.strip().split(",")
        assert H(base64.b64decode(m)) == base64.b64decode(pt)
    if cid == 0:
        a = 9121282642809701931333593728297233225556711250127745709186816755779879923737
        got = H(a.to_bytes(32, "big"))
        assert int.from_bytes(got[:32], "big") == 11423386531623885114587219621463106117140760157404497425836076043015227528156
        assert int.from_bytes(got[32:], "big") == 20262289731964024720969923714809935701428881933342918937283877214228227624643
    else:
        got = H(b"")
        assert int.from_bytes(got[:48], "big") == 315124130825307604287835216317628428134609737854237653839182597515996444073032649481416725367158979153513345579672
    rng = random.Random(17)
    for ln in (0, 1, 31, 32, 123, 124, 127, 128, 135, 136, 137, 200, 300):
        m = bytes(rng.randrange(256) for _ in range(ln))
        assert H(m) == c.marshal_g1(c.hash_to_g1(m)), ln
    if cid == 1:   # throughput form: one cofactor multiplication for both halves (h Q0 + h Q1 = h (Q0 + Q1))
        E.emu_hash_to_g1_shared_cofactor.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_char_p]
        msgs = [base64.b64decode(line.split(",")[0]) for line in open(os.path.join(HERE, "golden", fname))] + [b"", b"x" * 200]
        for m in msgs:
            out = ctypes.create_string_buffer(2 * nb)
            E.emu_hash_to_g1_shared_cofactor(m, len(m), out)
            assert out.raw == H(m)


@pytest.mark.parametrize("cid,c", CURVES)
def test_compressed_wire_formats(E, cid, c):
    """codec.cuh on the host against the oracle restatement of curves/altbn128.go:81-89,203-221,296-376 (altbn128)
    and of the zcash format the upstream bls12 library implements: round trips, both sign choices, infinity,
    non-residue abscissas, unreduced coordinates, malformed flags, subgroup check."""
    E.emu_compress.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_char_p, ctypes.c_char_p]
    E.emu_decompress.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_char_p, ctypes.c_int, ctypes.c_char_p]
    nb = c.nbytes
    rng = random.Random(31 + cid)

    def comp(group, rec):
        out = ctypes.create_string_buffer(group * nb)
        E.emu_compress(cid, group, rec, out)
        return out.raw

    def decomp(group, data, sub=0):
        out = ctypes.create_string_buffer(2 * group * nb)
        ok = E.emu_decompress(cid, group, data, sub, out)
        return out.raw, bool(ok)
    for _ in range(6):
        P = c.g1_mul(c.g1, rng.randrange(1, c.r))
        Q = c.g2_mul(c.g2, rng.randrange(1, c.r))
        for pt in (P, c.g1_neg(P)):
            assert comp(1, c.marshal_g1(pt)) == c.compress_g1(pt)
            assert decomp(1, c.compress_g1(pt), 1) == (c.marshal_g1(pt), True)
        for pt in (Q, c.g2_neg(Q)):
            assert comp(2, c.marshal_g2(pt)) == c.compress_g2(pt)
            assert decomp(2, c.compress_g2(pt)) == (c.marshal_g2(pt), True)
    assert decomp(2, c.compress_g2(c.g2), 1) == (c.marshal_g2(c.g2), True)
    assert comp(1, c.marshal_g1(None)) == c.compress_g1(None) and decomp(1, c.compress_g1(None)) == (bytes(2 * nb), True)
    assert comp(2, c.marshal_g2(None)) == c.compress_g2(None) and decomp(2, c.compress_g2(None)) == (bytes(4 * nb), True)
    flag = 0x80 if cid == 1 else 0
    seen = {True: 0, False: 0}
    for x in list(range(1, 30)) + [c.p - 1, c.p, c.p + 1]:
        for group in (1, 2):
            d = bytearray(bytes((group - 1) * nb) + (x % (1 << (8 * nb - 3))).to_bytes(nb, "big"))
            d[0] |= flag
            want_pt, want_ok = (c.decompress_g1 if group == 1 else c.decompress_g2)(bytes(d))
            rec, ok = decomp(group, bytes(d))
            assert ok == want_ok, (group, x)
            marshal = c.marshal_g1 if group == 1 else c.marshal_g2
            assert rec == (marshal(want_pt) if want_ok else bytes(2 * group * nb))
            seen[ok] += 1
            if ok and group == 1 and cid == 1:
                # bls12-381 G1 has a large cofactor: a random curve point is not in the subgroup
                assert decomp(1, bytes(d), 1)[1] == c.in_subgroup_g1(want_pt)
    assert seen[True] > 5 and seen[False] > 5
    if cid == 1:   # malformed flags: not marked compressed; infinity with a non-zero body
        d = bytearray(c.compress_g1(c.g1))
        d[0] &= 0x7F
        assert decomp(1, bytes(d)) == (bytes(2 * nb), False)
        d = bytearray(c.compress_g1(None))
        d[-1] = 1
        assert decomp(1, bytes(d)) == (bytes(2 * nb), False)
