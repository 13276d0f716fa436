"""The six-lane point aggregation (bgls_b200/csrc/agg.cuh: complete projective additions on plain-limb inputs, block
tree, binary inversion) run lane by lane on the host (tests/host_emul/emul_agg.cpp, test scaffolding) against the
oracle's AggregatePoints (reference: curves/curve.go:73-121)."""
import ctypes
import os
import random
import subprocess

import pytest

from oracle import c_oracle as C
from parity_util import CURVES, rand_points, scalars_bytes

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def emu():
    d = os.path.join(HERE, "host_emul")
    so, src = os.path.join(d, "libemul_agg.so"), os.path.join(d, "emul_agg.cpp")
    csrc = os.path.join(os.path.dirname(HERE), "bgls_b200", "csrc")
    newest = max(os.path.getmtime(os.path.join(csrc, f)) for f in ("agg.cuh", "sat.cuh", "slotvm.cuh", "arith.cuh"))
    if not os.path.exists(so) or os.path.getmtime(so) < max(newest, os.path.getmtime(src)):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", so, src])
    L = ctypes.CDLL(so)
    L.emu_agg6.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_char_p, ctypes.c_size_t, ctypes.c_char_p]
    return L


def run(emu, cid, c, group, blob, n):
    rec = 2 * group * c.nbytes
    out = ctypes.create_string_buffer(rec)
    emu.emu_agg6(cid, group, blob, n, out)
    return out.raw


@pytest.mark.parametrize("cid,c", CURVES)
@pytest.mark.parametrize("group", [1, 2])
def test_sums_match_the_oracle(emu, cid, c, group):
    rng = random.Random(11 * cid + group)
    rec = 2 * group * c.nbytes
    for n in (1, 2, 3, 7, 8, 9, 21):
        blob, _ = rand_points(cid, c, group, n, rng, nthreads=4)
        assert run(emu, cid, c, group, blob, n) == C.aggregate(cid, group, blob, n)
    # infinity records inside, a doubling (the same point twice) and a cancelling pair
    blob, ks = rand_points(cid, c, group, 6, rng, nthreads=4)
    P = [blob[i * rec:(i + 1) * rec] for i in range(6)]
    gen = c.marshal_g1(c.g1) if group == 1 else c.marshal_g2(c.g2)
    neg0 = C.scale_points(cid, group, gen, scalars_bytes([c.r - ks[0]]), 1)
    for pts in ([P[0], bytes(rec), P[1]], [bytes(rec)], [bytes(rec), bytes(rec), P[2]], [P[0], P[0]], [P[3], P[3], P[3], P[4]],
                [P[0], neg0], [P[0], P[1], neg0], [P[0], neg0, bytes(rec)], [P[5]] * 9):
        b = b"".join(pts)
        assert run(emu, cid, c, group, b, len(pts)) == C.aggregate(cid, group, b, len(pts)), len(pts)


@pytest.mark.parametrize("curve,N", [(0, 8), (1, 12)])
def test_binary_inversion(emu, curve, N):
    p = CURVES[curve][1].p
    rng = random.Random(5)
    for a in [1, 2, p - 1, p - 2, (p - 1) // 2, 1 << 200] + [rng.randrange(1, p) for _ in range(200)]:
        inp = (ctypes.c_uint32 * N)(*[(a >> (32 * i)) & 0xFFFFFFFF for i in range(N)])
        out = (ctypes.c_uint32 * N)()
        emu.emu_inv_plain(curve, inp, out)
        assert sum(w << (32 * i) for i, w in enumerate(out)) == pow(a, -1, p)


@pytest.mark.parametrize("curve,N", [(0, 8), (1, 12)])
def test_jacobi_symbol(emu, curve, N):
    """inv.cuh mp_jacobi against Euler's criterion: the residuosity test of hash-to-G1."""
    p = CURVES[curve][1].p
    rng = random.Random(6)
    for a in [0, 1, 2, 3, 4, p - 1, p - 2, (p - 1) // 2, 1 << 200, (1 << 200) + 1] + [rng.randrange(1, p) for _ in range(400)]:
        inp = (ctypes.c_uint32 * N)(*[(a >> (32 * i)) & 0xFFFFFFFF for i in range(N)])
        e = pow(a, (p - 1) // 2, p)
        assert emu.emu_jacobi(curve, inp) == (0 if a == 0 else 1 if e == 1 else -1), a
