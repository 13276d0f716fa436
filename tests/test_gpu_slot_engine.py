"""Slot engine (bgls_b200/csrc/slotvm.cuh: saturated limbs, lazy reduction, 8 lanes per group of 2 pairs) through the C ABI against the
oracle: the engine behind `PairingProduct` (/root/reference/curves/curve.go:125-170) whenever several products are in
flight or one product is large.  BGLS_ENGINE=slot forces it at every size."""
import os
import random
import threading

import pytest

from oracle import c_oracle as C
from parity_util import CURVES, make_aggregate, rand_points

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sctx():
    import bgls_b200
    old = os.environ.get("BGLS_ENGINE")
    os.environ["BGLS_ENGINE"] = "slot"
    try:
        c = bgls_b200.Context(0)
    finally:
        if old is None:
            del os.environ["BGLS_ENGINE"]
        else:
            os.environ["BGLS_ENGINE"] = old
    yield c
    c.close()


@pytest.mark.parametrize("cid,c", CURVES)
@pytest.mark.parametrize("n", [1, 2, 7, 8, 9, 16, 17, 33, 100, 1025])
def test_slot_product_bit_exact(sctx, cid, c, n):
    """GT bytes of random (rejecting) products equal the oracle's: ragged block sizes, one to many blocks."""
    rng = random.Random(7000 * cid + n)
    g1, _ = rand_points(cid, c, 1, n, rng)
    g2, _ = rand_points(cid, c, 2, n, rng)
    gt, one = sctx.pairing_product(cid, g1, g2, n)
    assert not one
    assert gt == C.pairing_product(cid, g1, g2, n, 8, 0)


@pytest.mark.parametrize("cid,c", CURVES)
def test_slot_aggregate_and_edges(sctx, cid, c):
    nb = c.nbytes
    rng = random.Random(31 + cid)
    g1, g2 = make_aggregate(cid, c, 40, rng)
    gt, ok = sctx.pairing_product(cid, g1, g2, 41)
    assert ok and gt == c.marshal_gt(c.fp12_one)
    # a pair with a point at infinity contributes 1 (curves/altbn128.go:478, bls12_381.go:341): G1 side, then G2 side
    h1 = g1[:2 * nb * 3] + bytes(2 * nb) + g1[2 * nb * 4:]
    gt, ok = sctx.pairing_product(cid, h1, g2, 41)
    assert gt == C.pairing_product(cid, h1, g2, 41, 8, 0) and not ok
    h2 = g2[:4 * nb * 5] + bytes(4 * nb) + g2[4 * nb * 6:]
    gt, ok = sctx.pairing_product(cid, g1, h2, 41)
    assert gt == C.pairing_product(cid, g1, h2, 41, 8, 0) and not ok
    # raw Miller products of two shards multiply to the full product (the NAF loop's raw value differs from the
    # oracle's binary loop by factors the final exponentiation kills, so only the exponentiated value is compared)
    parts = sctx.miller_product(cid, g1[:2 * nb * 20], g2[:4 * nb * 20], 20) + sctx.miller_product(cid, g1[2 * nb * 20:], g2[4 * nb * 20:], 21)
    gt, ok = sctx.final_exp_product(cid, parts, 2)
    assert ok and gt == c.marshal_gt(c.fp12_one)
    assert C.fp12_product(cid, parts, 2, True) == c.marshal_gt(c.fp12_one)


@pytest.mark.parametrize("cid,c", CURVES)
def test_auto_engine_switches_under_load(cid, c):
    """Default context: with many host threads inside the library the slot engine takes over (load estimate); the
    verdicts and GT bytes are the oracle's on every path."""
    import bgls_b200
    ctx = bgls_b200.Context(0)
    rng = random.Random(77 + cid)
    n = 65
    cases = []
    for k in range(6):
        g1, g2 = make_aggregate(cid, c, n - 1, rng)
        if k % 2:
            g1 = g1[:2 * c.nbytes] * 2 + g1[4 * c.nbytes:]
        cases.append((g1, g2, C.pairing_product(cid, g1, g2, n, 8, 0)))
    errs = []

    def work(i):
        for rep in range(6):
            g1, g2, exp = cases[(i + rep) % len(cases)]
            gt, ok = ctx.pairing_product(cid, g1, g2, n)
            if gt != exp or ok != (exp == c.marshal_gt(c.fp12_one)):
                errs.append((i, rep))
    ts = [threading.Thread(target=work, args=(i,)) for i in range(8)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errs
    ctx.close()


@pytest.mark.parametrize("cid,c", CURVES)
def test_slot_batch_ragged(sctx, cid, c):
    """Batch mode of the slot engine (one launch for all products: plan kernel, per-product trees and final
    exponentiations): ragged products -- empty, 1 pair, one block, several blocks -- with valid and rejecting ones mixed;
    every verdict is the oracle's."""
    rng = random.Random(555 + cid)
    nb = c.nbytes
    sizes = [0, 1, 2, 8, 9, 40, 0, 17, 64, 65, 3]
    g1s, g2s, exp = [], [], []
    for i, n in enumerate(sizes):
        if n == 0:
            g1, g2 = b"", b""
        elif n == 1:
            g1, g2 = c.marshal_g1(c.g1), c.marshal_g2(c.g2)
        else:
            g1, g2 = make_aggregate(cid, c, n - 1, rng, nthreads=8)
            if i % 3 == 0:   # corrupt: swap two G1 points
                g1 = g1[2 * nb:4 * nb] + g1[:2 * nb] + g1[4 * nb:]
        g1s.append(g1)
        g2s.append(g2)
        exp.append(True if n == 0 else C.pairing_product(cid, g1, g2, n, 8, 0) == c.marshal_gt(c.fp12_one))
    offs = [0]
    for n in sizes:
        offs.append(offs[-1] + n)
    assert sctx.pairing_check_batch(cid, b"".join(g1s), b"".join(g2s), offs) == exp
    assert any(exp) and not all(exp)


@pytest.mark.parametrize("cid,c", CURVES)
def test_slot_finish_of_gathered_partials(sctx, cid, c):
    """The finishing step of a sharded verification on the slot engine (k_slot_finish_bytes: product of the k gathered
    wire-form partials + final exponentiation in one launch) against the oracle's product of the same bytes, for every
    k one block covers (1..8) and beyond (9: the machine's import / tree / finish path)."""
    rng = random.Random(900 + cid)
    nb = c.nbytes
    parts = b""
    for k in range(1, 10):
        g1, _ = rand_points(cid, c, 1, 3, rng)
        g2, _ = rand_points(cid, c, 2, 3, rng)
        parts += sctx.miller_product(cid, g1, g2, 3)
        assert len(parts) == 12 * nb * k
        gt, one = sctx.final_exp_product(cid, parts, k)
        assert not one and gt == C.fp12_product(cid, parts, k, True), k
    g1, g2 = make_aggregate(cid, c, 23, rng)
    cut = [0, 5, 6, 13, 24]
    shards = b"".join(sctx.miller_product(cid, g1[2 * nb * a:2 * nb * b], g2[4 * nb * a:4 * nb * b], b - a) for a, b in zip(cut, cut[1:]))
    gt, one = sctx.final_exp_product(cid, shards, 4)
    assert one and gt == c.marshal_gt(c.fp12_one)
