"""BASELINE.json-size cases on the GPU through size-independent properties (the oracle would take
minutes at these sizes): inputs are produced by the engine's own ScalePoints kernel with known
discrete logs, spot-checked against the oracle, and the verdicts / sums must follow."""
import random

import pytest

from oracle import c_oracle as C
from parity_util import CURVES, scalars_bytes

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import bgls_b200
    c = bgls_b200.Context(0)
    yield c
    c.close()


def synth_aggregate(ctx, cid, c, n, seed):
    """(g1, g2) of a valid n-signer aggregate (n+1 pairs) built on the GPU: H_i = h_i G1, pk_i = s_i G2."""
    rng = random.Random(seed)
    hs = [rng.randrange(1, c.r) for _ in range(n)]
    ss = [rng.randrange(1, c.r) for _ in range(n)]
    g1 = ctx.scale_points(cid, 1, c.marshal_g1(c.g1) * n, scalars_bytes(hs), n)
    g2 = ctx.scale_points(cid, 2, c.marshal_g2(c.g2) * n, scalars_bytes(ss), n)
    tot = sum(h * s for h, s in zip(hs, ss)) % c.r
    neg = ctx.scale_points(cid, 1, c.marshal_g1(c.g1), scalars_bytes([(c.r - tot) % c.r]), 1)
    # spot-check the generated points against the oracle
    nb = c.nbytes
    for i in (0, n // 2, n - 1):
        assert g1[2 * nb * i:2 * nb * (i + 1)] == C.scale_points(cid, 1, c.marshal_g1(c.g1), scalars_bytes([hs[i]]), 1)
        assert g2[4 * nb * i:4 * nb * (i + 1)] == C.scale_points(cid, 2, c.marshal_g2(c.g2), scalars_bytes([ss[i]]), 1)
    return g1 + neg, g2 + c.marshal_g2(c.g2)


def test_config3_multisig_65536_g2_keys(ctx):
    """config 3: bls12-381 multi-signature verify, 65,536 signers: AggregatePoints over 12 MiB of G2 keys
    (bit-exact against the oracle sum and against (sum k_i) G2), then the 2-pairing check."""
    cid, c = CURVES[1]
    n = 65536
    rng = random.Random(3)
    ks = [rng.randrange(1, c.r) for _ in range(n)]
    keys = ctx.scale_points(cid, 2, c.marshal_g2(c.g2) * n, scalars_bytes(ks), n)
    agg = ctx.aggregate_points(cid, 2, keys, n)
    ksum = sum(ks) % c.r
    assert agg == C.scale_points(cid, 2, c.marshal_g2(c.g2), scalars_bytes([ksum]), 1)
    assert agg == C.aggregate(cid, 2, keys, n, 8)
    # multi-signature on one message: sigma = (sum k_i) H ; e(-H, agg) e(sigma, g2) == 1  (bgls/bgls.go:65-70,89-92)
    h = rng.randrange(1, c.r)
    H = C.scale_points(cid, 1, c.marshal_g1(c.g1), scalars_bytes([h]), 1)
    negH = C.scale_points(cid, 1, c.marshal_g1(c.g1), scalars_bytes([(c.r - h) % c.r]), 1)
    sigma = C.scale_points(cid, 1, H, scalars_bytes([ksum]), 1)
    gt, ok = ctx.pairing_product(cid, negH + sigma, agg + c.marshal_g2(c.g2), 2)
    assert ok and gt == c.marshal_gt(c.fp12_one)
    _, ok = ctx.pairing_product(cid, negH + sigma, keys[:192] + c.marshal_g2(c.g2), 2)
    assert not ok


@pytest.mark.parametrize("cid,c,n", [(0, CURVES[0][1], 40000), (1, CURVES[1][1], 131072)])
def test_large_aggregate_verify_and_sharding(ctx, cid, c, n):
    """config 4 per-GPU share (bls12-381, 2^20 / 8 = 131,072 signers) and a large altbn128 product: valid
    aggregate -> identity; one swapped message -> reject; product of two shard Miller products == unsharded."""
    g1, g2 = synth_aggregate(ctx, cid, c, n, 1000 + cid)
    gt, ok = ctx.pairing_product(cid, g1, g2, n + 1)
    assert ok and gt == c.marshal_gt(c.fp12_one)
    nb = c.nbytes
    bad = g1[2 * nb:4 * nb] + g1[2 * nb:]
    gt_bad, ok = ctx.pairing_product(cid, bad, g2, n + 1)
    assert not ok
    half = (n + 1) // 2
    parts = ctx.miller_product(cid, g1[:2 * nb * half], g2[:4 * nb * half], half) + \
        ctx.miller_product(cid, g1[2 * nb * half:], g2[4 * nb * half:], n + 1 - half)
    gt2, ok2 = ctx.final_exp_product(cid, parts, 2)
    assert ok2 and gt2 == gt
    # the rejected product is still a deterministic GT element: both engines' regimes must agree on it
    small = 4000
    gts, _ = ctx.pairing_product(cid, bad[:2 * nb * small], g2[:4 * nb * small], small)
    assert gts == C.pairing_product(cid, bad[:2 * nb * small], g2[:4 * nb * small], small, 8, 0)


def test_config5_batch_of_independent_verifies(ctx):
    """config 5 shape (independent altbn128 aggregate verifies, throughput mode): 512 checks x 65 pairs with a
    known pattern of corrupted signatures."""
    cid, c = CURVES[0]
    nb = c.nbytes
    signers, nbatch = 64, 512
    g1, g2 = synth_aggregate(ctx, cid, c, signers, 77)
    G1, G2, offsets, expect = [], [], [0], []
    for b in range(nbatch):
        good = (b % 7) != 3
        a1 = g1 if good else g1[:-2 * nb] + c.marshal_g1(c.g1)
        G1.append(a1)
        G2.append(g2)
        offsets.append(offsets[-1] + signers + 1)
        expect.append(good)
    assert ctx.pairing_check_batch(cid, b"".join(G1), b"".join(G2), offsets) == expect


def test_batch_shared_accumulator_ragged(ctx):
    """Throughput path of pairing_check_batch (one block per check, shared Miller accumulators): ragged check
    sizes around the per-thread grouping boundaries, valid and corrupted, against per-check PairingProduct."""
    cid, c = CURVES[0]
    nb = c.nbytes
    rng = random.Random(91)
    sizes = [16, 17, 63, 64, 65, 127, 128, 129, 200, 257, 511, 512, 513, 600, 40, 33] * 3
    G1, G2, offsets, expect = [], [], [0], []
    for i, n in enumerate(sizes):
        g1, g2 = synth_aggregate(ctx, cid, c, n - 1, 500 + n + 1000 * (i // 16))
        good = i % 3 != 1
        if not good:   # swap two hashes: still on the curve, product no longer 1
            g1 = g1[2 * nb:4 * nb] + g1[:2 * nb] + g1[4 * nb:] if n > 2 else g1[:-2 * nb] + c.marshal_g1(c.g1)
        _, ok = ctx.pairing_product(cid, g1, g2, n)
        assert ok == good
        G1.append(g1)
        G2.append(g2)
        offsets.append(offsets[-1] + n)
        expect.append(good)
    assert offsets[-1] >= 8192
    assert ctx.pairing_check_batch(cid, b"".join(G1), b"".join(G2), offsets) == expect


def _engine_ctx(name):
    import os

    import bgls_b200
    old = os.environ.get("BGLS_ENGINE")
    os.environ["BGLS_ENGINE"] = name
    try:
        return bgls_b200.Context(0)
    finally:
        if old is None:
            del os.environ["BGLS_ENGINE"]
        else:
            os.environ["BGLS_ENGINE"] = old


@pytest.mark.parametrize("cid,c", CURVES)
def test_large_rejecting_product_bytes_on_every_engine(ctx, cid, c):
    """GT bytes of a REJECTING 40,001-pair product against the C oracle (shared final exponentiation, a second or two on
    the host cores), on the default path (slot engine at this size) and on the thread engine's shared-accumulator kernel
    (k_miller_product_shared, >= 32,768 pairs) -- and of a 4,000-pair product on all four engine settings."""
    g1, g2 = synth_aggregate(ctx, cid, c, 2000, 11 + cid)
    nb = c.nbytes
    # 20 tiles of the 2,000 signer pairs + the (-sigma, g2) pair once: a product that is NOT the identity
    big1, big2, n = g1[:2 * nb * 2000] * 20 + g1[2 * nb * 2000:], g2[:4 * nb * 2000] * 20 + g2[4 * nb * 2000:], 40001
    want = C.pairing_product(cid, big1, big2, n, 16, 0)
    assert want != c.marshal_gt(c.fp12_one)
    gt, ok = ctx.pairing_product(cid, big1, big2, n)
    assert gt == want and not ok
    small1, small2, m = g1[:2 * nb * 2000] * 2, g2[:4 * nb * 2000] * 2, 4000
    want_small = C.pairing_product(cid, small1, small2, m, 16, 0)
    assert ctx.pairing_product(cid, small1, small2, m)[0] == want_small
    for name in ("thread", "machine", "slot"):
        e = _engine_ctx(name)
        try:
            assert e.pairing_product(cid, small1, small2, m)[0] == want_small, name
            if name != "machine":   # (the machine at 40,001 pairs is covered by the property tests above)
                gt, ok = e.pairing_product(cid, big1, big2, n)
                assert gt == want and not ok, name
        finally:
            e.close()


@pytest.mark.parametrize("cid,c", CURVES)
def test_batch_path_verdicts_against_oracle(ctx, cid, c):
    """Config-5 shape at reduced count: 40 checks x 257 pairs (the throughput batch path: >= 8,192 pairs in total), some
    corrupted; every verdict must be the oracle's for that check's own product."""
    g1, g2 = synth_aggregate(ctx, cid, c, 256, 23 + cid)
    nb = c.nbytes
    bad1 = g1[2 * nb:4 * nb] + g1[:2 * nb] + g1[4 * nb:]          # two messages swapped
    bad2 = g2[:4 * nb * 5] + g2[4 * nb * 6:4 * nb * 7] + g2[4 * nb * 6:]   # a key replaced by its neighbour
    variants = [(g1, g2), (bad1, g2), (g1, bad2)]
    exp = [C.pairing_product(cid, a, b, 257, 16, 0) == c.marshal_gt(c.fp12_one) for a, b in variants]
    assert exp == [True, False, False]
    order = [0, 1, 0, 0, 2, 0, 1, 0] * 5
    all1 = b"".join(variants[k][0] for k in order)
    all2 = b"".join(variants[k][1] for k in order)
    offs = [257 * i for i in range(len(order) + 1)]
    assert ctx.pairing_check_batch(cid, all1, all2, offs) == [exp[k] for k in order]
