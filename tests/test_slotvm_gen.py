"""Slot engine on the CPU: (1) the generated programs evaluated over big integers (tools/gen_slotvm.py: Emu) against
the Python oracle; (2) the C++ interpreter itself (bgls_b200/csrc/slotvm.cuh + sat.cuh), compiled for the host with
the carry primitives emulated, against the C oracle -- the same code the kernel runs, lane by lane."""
import ctypes
import os
import random
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))

import gen_slotvm as G  # noqa: E402
from oracle import bgls_oracle as O  # noqa: E402
from oracle import c_oracle as C  # noqa: E402
from parity_util import CURVES, make_aggregate  # noqa: E402

CFG = {0: G.BN, 1: G.BLS}


@pytest.fixture(scope="module")
def emu_lib():
    d = os.path.join(ROOT, "tests", "host_emul")
    so, src = os.path.join(d, "libemul_slot.so"), os.path.join(d, "emul_slot.cpp")
    deps = [src] + [os.path.join(ROOT, "bgls_b200", "csrc", f) for f in ("slotvm.cuh", "sat.cuh", "slotvm_tables.cuh", "arith.cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(x) > os.path.getmtime(so) for x in deps):
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-o", so, src])
    return ctypes.CDLL(so)


def test_tables_are_current():
    """slotvm_tables.cuh is what tools/gen_slotvm.py emits (a stale table would silently run an old schedule)."""
    import tempfile
    with tempfile.TemporaryDirectory() as t:
        p = os.path.join(t, "t.cuh")
        G.emit(p)
        assert open(p).read() == open(os.path.join(ROOT, "bgls_b200", "csrc", "slotvm_tables.cuh")).read()


@pytest.mark.parametrize("cid,c", CURVES)
@pytest.mark.parametrize("g,k", G.SHAPES)
def test_programs_over_big_integers(cid, c, g, k):
    eng = G.engine(CFG[cid], g, k)
    emu = G.Emu(eng)
    rng = random.Random(5 + cid)
    for trial in range(3 if k > 1 else 1):
        Ps = [c.g1_mul(c.g1, rng.randrange(1, c.r)) for _ in range(k)]
        Qs = [c.g2_mul(c.g2, rng.randrange(1, c.r)) for _ in range(k)]
        if trial == 1:
            Ps[k - 1] = None      # a point at infinity in a group that shares its accumulator: the pair contributes 1
        if trial == 2:
            Qs[0] = None
        g1 = b"".join(c.marshal_g1(P) for P in Ps)
        g2 = b"".join(c.marshal_g2(Q) for Q in Qs)
        raw_c = C.miller_product(cid, g1, g2, k)
        if k == 1:
            # binary loop: the raw Miller value is the oracle's, bit for bit (same formulas as oracle/pairing_impl.h)
            assert c.marshal_gt(emu.miller(Ps, Qs, use_naf=False)) == raw_c
        # released sequence (NAF on altbn128; shared accumulator for k > 1): equal after the final exponentiation
        for naf in (False, True):
            assert C.fp12_product(cid, c.marshal_gt(emu.miller(Ps, Qs, use_naf=naf)), 1, True) == C.fp12_product(cid, raw_c, 1, True)
    for name, s in eng.progs.items():
        G.check_hazards(s.rounds, g)
        assert s.nslots <= eng.nslots < G.CONST0


@pytest.mark.parametrize("cid,c", CURVES)
@pytest.mark.parametrize("g,k", G.SHAPES)
def test_interpreter_on_host(emu_lib, cid, c, g, k):
    nb = c.nbytes
    rng = random.Random(100 + cid)
    for n in (1, 3, 5, 8, 8 * k - 1):
        g1, g2 = make_aggregate(cid, c, n - 1, rng, nthreads=4) if n > 1 else (c.marshal_g1(c.g1), c.marshal_g2(c.g2))
        if n == 5:   # an infinity pair inside (G1 side), and one on the G2 side
            g1 = g1[:2 * nb] + bytes(2 * nb) + g1[4 * nb:]
            g2 = g2[:4 * nb * 2] + bytes(4 * nb) + g2[4 * nb * 3:]
        out = ctypes.create_string_buffer(12 * nb)
        assert emu_lib.emu_slot_miller_product(cid, g, k, g1, g2, n, out) == 0
        assert C.fp12_product(cid, out.raw, 1, True) == C.pairing_product(cid, g1, g2, n, 4, 0)


@pytest.mark.parametrize("curve,N,p", [(0, 8, G.BN_P), (1, 12, G.BLS_P)])
def test_saturated_primitives(emu_lib, curve, N, p):
    """sat.cuh: product, reduction, Fp2 multiplication / squaring / xi on edge and random operands."""
    R = 1 << (32 * N)
    Ri = pow(R, -1, p)
    xa = 9 if curve == 0 else 1
    rng = random.Random(1)
    words = lambda v: [(v >> (32 * i)) & 0xFFFFFFFF for i in range(N)]
    val = lambda ws: sum(w << (32 * i) for i, w in enumerate(ws))
    for it in range(400):
        vals = [rng.choice([0, 1, p - 1, p - 2, (p - 1) // 2, R % p]) for _ in range(4)] if it < 60 else [rng.randrange(p) for _ in range(4)]
        a0, a1, b0, b1 = vals
        inp = (ctypes.c_uint32 * (4 * N))(*sum((words(v) for v in vals), []))
        out = (ctypes.c_uint32 * (11 * N))()
        emu_lib.emu_sat_ops(curve, inp, out)
        o = list(out)
        got = [val(o[k * N:(k + 1) * N]) for k in range(9)]
        exp = [(a0 * b0 - a1 * b1) * Ri % p, (a0 * b1 + a1 * b0) * Ri % p, (a0 * a0 - a1 * a1) * Ri % p, 2 * a0 * a1 * Ri % p,
               (xa * a0 - a1) % p, (xa * a1 + a0) % p, a0 * pow(2, -1, p) % p, (a0 + b1) % p, (a0 - b1) % p]
        assert got == exp, (curve, it)
        assert val(o[9 * N:11 * N]) == a0 * b0
