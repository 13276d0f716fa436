"""GPU parity tests proper: the CUDA path, called through the C ABI, against the oracle on the
same seeded inputs (bit-exact) and against the committed golden vectors."""
import random

import pytest

from oracle import c_oracle as C
from parity_util import CURVES, GOLD, make_aggregate, rand_points, scalars_bytes

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import bgls_b200
    c = bgls_b200.Context(0)
    yield c
    c.close()


@pytest.mark.parametrize("cid,c", CURVES)
def test_golden_vectors(ctx, cid, c):
    g = GOLD[c.name]
    g1, g2, n = bytes.fromhex(g["g1"]), bytes.fromhex(g["g2"]), g["n"]
    gt, one = ctx.pairing_product(cid, g1, g2, n)
    assert gt.hex() == g["product_gt"] and not one
    assert ctx.pair(cid, c.marshal_g1(c.g1), c.marshal_g2(c.g2)).hex() == g["gen_gt"]
    nb = c.nbytes
    assert ctx.pair(cid, g1[:2 * nb], g2[:4 * nb]).hex() == g["pair0_gt"]
    assert ctx.aggregate_points(cid, 1, g1, n).hex() == g["sum_g1"]
    assert ctx.aggregate_points(cid, 2, g2, n).hex() == g["sum_g2"]


@pytest.mark.parametrize("cid,c", CURVES)
def test_golden_aggregate_signature(ctx, cid, c):
    """bgls/bgls_test.go:40-77 shape on the golden 3-signer aggregate (real HashToG1 points)."""
    a = GOLD[c.name]["agg"]
    hs = [bytes.fromhex(h) for h in a["hashes"]]
    pks = [bytes.fromhex(h) for h in a["pubkeys"]]
    neg = c.marshal_g1(c.g1_neg(c.unmarshal_g1(bytes.fromhex(a["sig"]))))
    g2gen = bytes.fromhex(a["g2gen"])
    gt, ok = ctx.pairing_product(cid, b"".join(hs) + neg, b"".join(pks) + g2gen, 4)
    assert ok and gt == c.marshal_gt(c.fp12_one)
    # swapped messages, missing key, wrong signature -> reject
    _, ok = ctx.pairing_product(cid, hs[1] + hs[0] + hs[2] + neg, b"".join(pks) + g2gen, 4)
    assert not ok
    _, ok = ctx.pairing_product(cid, hs[0] + hs[1] + neg, pks[0] + pks[1] + g2gen, 3)
    assert not ok
    bad = c.marshal_g1(c.g1_neg(c.g1_add(c.unmarshal_g1(bytes.fromhex(a["sig"])), c.g1)))
    _, ok = ctx.pairing_product(cid, b"".join(hs) + bad, b"".join(pks) + g2gen, 4)
    assert not ok


@pytest.mark.parametrize("cid,c", CURVES)
@pytest.mark.parametrize("n", [1, 2, 5, 33, 100])
def test_pairing_product_random(ctx, cid, c, n):
    """curves/curve_test.go:143-165 (TestPairingProd): product == prod of pairs, bit-exact GT bytes."""
    rng = random.Random(1000 * cid + n)
    g1, _ = rand_points(cid, c, 1, n, rng)
    g2, _ = rand_points(cid, c, 2, n, rng)
    gt, one = ctx.pairing_product(cid, g1, g2, n)
    assert gt == C.pairing_product(cid, g1, g2, n, 8, 0)
    assert not one
    if n == 5:
        # product of individually exponentiated pairings (the reference's structure)
        assert gt == C.pairing_product(cid, g1, g2, n, 4, 1)
        nb = c.nbytes
        acc = c.marshal_gt(c.fp12_one)
        for i in range(n):
            acc = ctx.gt_mul(cid, acc, ctx.pair(cid, g1[2 * nb * i:2 * nb * (i + 1)], g2[4 * nb * i:4 * nb * (i + 1)]))
        assert acc == gt


@pytest.mark.parametrize("cid,c", CURVES)
def test_edge_cases(ctx, cid, c):
    one = c.marshal_gt(c.fp12_one)
    # n = 0: empty product is the identity
    gt, ok = ctx.pairing_product(cid, b"", b"", 0)
    assert gt == one and ok
    # Pair with infinity is the identity (altbn128.go:478, bls12_381.go:341)
    assert ctx.pair(cid, c.marshal_g1(None), c.marshal_g2(c.g2)) == one
    assert ctx.pair(cid, c.marshal_g1(c.g1), c.marshal_g2(None)) == one
    # infinity inside a product contributes nothing
    rng = random.Random(9)
    g1, _ = rand_points(cid, c, 1, 3, rng)
    g2, _ = rand_points(cid, c, 2, 3, rng)
    nb = c.nbytes
    g1i = g1[:2 * nb] + c.marshal_g1(None) + g1[2 * nb:]
    g2i = g2[:4 * nb] + c.marshal_g2(c.g2) + g2[4 * nb:]
    assert ctx.pairing_product(cid, g1i, g2i, 4)[0] == ctx.pairing_product(cid, g1, g2, 3)[0]
    # e(P,Q) * e(-P,Q) == 1
    P = c.unmarshal_g1(g1[:2 * nb])
    gt, ok = ctx.pairing_product(cid, g1[:2 * nb] + c.marshal_g1(c.g1_neg(P)), g2[:4 * nb] * 2, 2)
    assert ok and gt == one
    if cid == 1:  # bls12: 0x40 infinity flag accepted
        flagged = bytes([0x40]) + bytes(2 * nb - 1)
        assert ctx.pair(cid, flagged, c.marshal_g2(c.g2)) == one


@pytest.mark.parametrize("cid,c", CURVES)
def test_sharded_miller_product(ctx, cid, c):
    """per-shard Miller products, exchanged, multiplied, one final exponentiation."""
    rng = random.Random(77 + cid)
    n, shards = 37, 4
    g1, _ = rand_points(cid, c, 1, n, rng)
    g2, _ = rand_points(cid, c, 2, n, rng)
    nb = c.nbytes
    full, _ = ctx.pairing_product(cid, g1, g2, n)
    parts = b""
    for s in range(shards):
        lo, hi = n * s // shards, n * (s + 1) // shards
        parts += ctx.miller_product(cid, g1[2 * nb * lo:2 * nb * hi], g2[4 * nb * lo:4 * nb * hi], hi - lo)
    gt, _ = ctx.final_exp_product(cid, parts, shards)
    assert gt == full == C.pairing_product(cid, g1, g2, n, 8, 0)
    # the raw Miller product differs from the oracle's by subfield factors only (NAF loop, projective
    # scalings): the oracle's final exponentiation of it must give the same GT element
    assert C.fp12_product(cid, ctx.miller_product(cid, g1, g2, n), 1, True) == full


@pytest.mark.parametrize("cid,c", CURVES)
@pytest.mark.parametrize("group", [1, 2])
def test_aggregate_points(ctx, cid, c, group):
    """curves/curve_test.go:167-186 (TestAggregation) + ragged sizes, infinity, repeats, cancellation."""
    rng = random.Random(5 * cid + group)
    rec = 2 * group * c.nbytes
    for n in (1, 2, 3, 4, 6, 8, 31, 32, 33, 1000):
        pts, ks = rand_points(cid, c, group, n, rng)
        got = ctx.aggregate_points(cid, group, pts, n)
        assert got == C.aggregate(cid, group, pts, n, 8), n
        gen = c.marshal_g1(c.g1) if group == 1 else c.marshal_g2(c.g2)
        assert got == C.scale_points(cid, group, gen, scalars_bytes([sum(ks) % c.r]), 1), n
    pts, _ = rand_points(cid, c, group, 3, rng)
    neg = c.marshal_g1(c.g1_neg(c.unmarshal_g1(pts[:rec]))) if group == 1 else c.marshal_g2(c.g2_neg(c.unmarshal_g2(pts[:rec])))
    inf = bytes(rec)
    mix = pts[:rec] + inf + pts[:rec] + pts[rec:2 * rec] + neg + inf + pts[2 * rec:]
    assert ctx.aggregate_points(cid, group, mix, 7) == C.aggregate(cid, group, mix, 7, 1)
    assert ctx.aggregate_points(cid, group, pts[:rec] + neg, 2) == inf


@pytest.mark.parametrize("cid,c", CURVES)
@pytest.mark.parametrize("group", [1, 2])
def test_scale_points(ctx, cid, c, group):
    """curves/curve_test.go:188-208 (TestScaling) and :120-141 (TestMul: k=0, 1, r-1)."""
    rng = random.Random(50 * cid + group)
    n = 40
    pts, _ = rand_points(cid, c, group, n, rng)
    ks = [0, 1, c.r - 1, c.r, 2] + [rng.randrange(c.r) for _ in range(n - 5)]
    sc = scalars_bytes(ks)
    assert ctx.scale_points(cid, group, pts, sc, n) == C.scale_points(cid, group, pts, sc, n, 8)


@pytest.mark.parametrize("cid,c", CURVES)
def test_batch_check(ctx, cid, c):
    rng = random.Random(31 + cid)
    sizes = [3, 1, 0, 5, 2]
    g1 = g2 = b""
    offsets = [0]
    expect = []
    nb = c.nbytes
    for k, n in enumerate(sizes):
        a1, a2 = make_aggregate(cid, c, n, rng)
        good = (k % 2 == 0)
        if not good:  # corrupt: replace the signature pair's G1 by the generator
            a1 = a1[:-2 * nb] + c.marshal_g1(c.g1)
        g1 += a1
        g2 += a2
        offsets.append(offsets[-1] + n + 1)
        expect.append(good)
    assert ctx.pairing_check_batch(cid, g1, g2, offsets) == expect


@pytest.mark.parametrize("cid,c", CURVES)
def test_aggregate_verify_config_shape(ctx, cid, c):
    """BASELINE config 2 shape (1024 signers -> 1025 pairs; 257 for bls12 to bound CPU time)."""
    n = 1024 if cid == 0 else 256
    rng = random.Random(2024 + cid)
    g1, g2 = make_aggregate(cid, c, n, rng)
    gt, ok = ctx.pairing_product(cid, g1, g2, n + 1)
    assert ok and gt == c.marshal_gt(c.fp12_one)
    assert gt == C.pairing_product(cid, g1, g2, n + 1, 8, 0)
    # flip one hashed message -> reject, and GT bytes still match the oracle
    nb = c.nbytes
    bad = g1[2 * nb:4 * nb] + g1[2 * nb:]
    gt, ok = ctx.pairing_product(cid, bad, g2, n + 1)
    assert not ok and gt == C.pairing_product(cid, bad, g2, n + 1, 8, 0)


@pytest.mark.parametrize("cid,c", CURVES)
def test_hash_to_g1_reference_vectors_gpu(ctx, cid, c):
    """curves/curve_test.go:210-244 (TestG1HashVectors), curves/altbn128_test.go:13-24, curves/bls12_test.go:57-67
    through bgls_hash_to_g1, plus ragged lengths against the oracle in one batched call."""
    import base64
    import os
    fname = "altbn128G1Hash.dat" if cid == 0 else "bls12G1Hash.dat"
    msgs, want = [], []
    for line in open(os.path.join(os.path.dirname(__file__), "golden", fname)):
        m, pt = line.strip().split(",")
        msgs.append(base64.b64decode(m))
        want.append(base64.b64decode(pt))
    rng = random.Random(4)
    for ln in (0, 1, 31, 32, 64, 127, 128, 135, 136, 137, 255, 300):
        m = bytes(rng.randrange(256) for _ in range(ln))
        msgs.append(m)
        want.append(c.marshal_g1(c.hash_to_g1(m)))
    if cid == 0:
        a = 9121282642809701931333593728297233225556711250127745709186816755779879923737
        msgs.append(a.to_bytes(32, "big"))
        want.append((11423386531623885114587219621463106117140760157404497425836076043015227528156).to_bytes(32, "big") +
                    (20262289731964024720969923714809935701428881933342918937283877214228227624643).to_bytes(32, "big"))
    got = ctx.hash_to_g1(cid, msgs)
    rec = 2 * c.nbytes
    for i, w in enumerate(want):
        assert got[i * rec:(i + 1) * rec] == w, i


def test_bls12_throughput_hash_equals_latency_hash(ctx):
    """bls12-381 HashToG1 in its throughput form (k_hash_to_g1_bls_one: one cofactor multiplication per message) against
    the two-lanes-per-message form and the reference vectors."""
    import base64
    import os
    import bgls_b200
    cid, c = CURVES[1]
    old = os.environ.get("BGLS_HASH")
    os.environ["BGLS_HASH"] = "pool"
    try:
        pctx = bgls_b200.Context(0)
    finally:
        if old is None:
            del os.environ["BGLS_HASH"]
        else:
            os.environ["BGLS_HASH"] = old
    try:
        msgs, want = [], []
        for line in open(os.path.join(os.path.dirname(__file__), "golden", "bls12G1Hash.dat")):
            m, pt = line.strip().split(",")
            msgs.append(base64.b64decode(m))
            want.append(base64.b64decode(pt))
        rng = random.Random(43)
        msgs += [bytes(rng.randrange(256) for _ in range(rng.choice((0, 1, 31, 64, 124, 128, 200)))) for _ in range(300)]
        got = pctx.hash_to_g1(cid, msgs)
        for i, w in enumerate(want):
            assert got[96 * i:96 * (i + 1)] == w, i
        assert got == ctx.hash_to_g1(cid, msgs)
    finally:
        pctx.close()


def test_pooled_try_and_increment_is_the_sequential_loop():
    """altbn128 HashToG1 in its throughput form (k_hash_to_g1_bn_pool: a warp re-deals its lanes over the messages
    still open every round) finds the counter the sequential loop of curves/hash.go:53-77 finds: reference vectors,
    ragged message counts (partial warps) and 2,000 random messages against the oracle."""
    import base64
    import os
    import bgls_b200
    cid, c = CURVES[0]
    old = os.environ.get("BGLS_HASH")
    os.environ["BGLS_HASH"] = "pool"
    try:
        pctx = bgls_b200.Context(0)
    finally:
        if old is None:
            del os.environ["BGLS_HASH"]
        else:
            os.environ["BGLS_HASH"] = old
    try:
        msgs, want = [], []
        for line in open(os.path.join(os.path.dirname(__file__), "golden", "altbn128G1Hash.dat")):
            m, pt = line.strip().split(",")
            msgs.append(base64.b64decode(m))
            want.append(base64.b64decode(pt))
        rng = random.Random(41)
        for k in range(2000):
            m = bytes(rng.randrange(256) for _ in range(rng.choice((0, 1, 31, 32, 33, 135, 136, 200))))
            msgs.append(m)
            want.append(C.hash_to_g1(cid, m) if hasattr(C, "hash_to_g1") else c.marshal_g1(c.hash_to_g1(m)))
        for n in (1, 5, 31, 32, 33, 64, len(msgs)):
            got = pctx.hash_to_g1(cid, msgs[:n])
            for i in range(n):
                assert got[i * 64:(i + 1) * 64] == want[i], (n, i)
    finally:
        pctx.close()


def test_scheme_layer_end_to_end():
    """bgls/bgls_test.go:19-77 (TestSingleSigner, TestAggregation) and bgls/blsKosk_test.go multisig shape, written
    against the engine-backed mirror of the reference API: every hash, scalar multiplication, point sum and pairing
    runs on the GPU."""
    from bgls_b200 import bgls
    from bgls_b200.curves import Altbn128, Bls12
    rng = random.Random(2018)
    for curve in (Altbn128, Bls12):
        sk, vk, err = bgls.KeyGen(curve, rng)
        assert err is None
        d = bytes(rng.randrange(256) for _ in range(64))
        sig = bgls.Sign(curve, sk, d)
        assert bgls.VerifySingleSignature(curve, sig, vk, d)
        sig2, _ = sig.Copy().Add(curve.GetG1())
        assert not bgls.VerifySingleSignature(curve, sig2, vk, d)
        N = 6
        msgs = [bytes(rng.randrange(256) for _ in range(32)) for _ in range(N)]
        keys = [bgls.KeyGen(curve, rng) for _ in range(N)]
        sigs = [bgls.Sign(curve, k[0], m) for k, m in zip(keys, msgs)]
        pubs = [k[1] for k in keys]
        agg = bgls.AggregateSignatures(sigs)
        assert bgls.VerifyAggregateSignature(curve, agg, pubs, msgs)
        assert not bgls.VerifyAggregateSignature(curve, agg, pubs[:N - 1], msgs)
        skf, vkf, _ = bgls.KeyGen(curve, rng)
        agg2 = bgls.AggregateSignatures(sigs + [bgls.Sign(curve, skf, msgs[0])])
        assert not bgls.VerifyAggregateSignature(curve, agg2, pubs + [vkf], msgs + [msgs[0]])   # duplicate message
        assert not bgls.VerifyAggregateSignature(curve, agg2, pubs, msgs)                        # wrong signature
        assert not bgls.VerifyAggregateSignature(curve, agg, pubs, [msgs[1], msgs[0]] + msgs[2:])  # swapped
        # Kosk multi-signature: 8 signers, one message (bgls/blsKosk_test.go:35-64)
        m = bytes(rng.randrange(256) for _ in range(64))
        ks = [bgls.KeyGen(curve, rng) for _ in range(8)]
        msig = bgls.AggregateSignatures([bgls.KoskSign(curve, k[0], m) for k in ks])
        assert bgls.KoskVerifyMultiSignature(curve, msig, [k[1] for k in ks], m)
        assert not bgls.KoskVerifyMultiSignature(curve, msig, [k[1] for k in ks[:-1]], m)
        # batch multi-signature (bgls/blsKosk.go:126-133), multiplicity (:137-150), distinct messages
        ms = [bytes(rng.randrange(256) for _ in range(16)) for _ in range(3)]
        groups = [[bgls.KeyGen(curve, rng) for _ in range(3)] for _ in ms]
        msigs = [bgls.AggregateSignatures([bgls.KoskSign(curve, k[0], mm) for k in g]) for g, mm in zip(groups, ms)]
        assert bgls.KoskVerifyBatchMultiSignature(curve, msigs, [[k[1] for k in g] for g in groups], ms)
        assert not bgls.KoskVerifyBatchMultiSignature(curve, msigs, [[k[1] for k in g] for g in groups], ms[::-1])
        mult = [1, 3, 2, 1, 1, 5, 1, 2]
        msig_m = bgls.AggregateSignatures([bgls.KoskSign(curve, k[0], m).Mul(f) for k, f in zip(ks, mult)])
        assert bgls.KoskVerifyMultiSignatureWithMultiplicity(curve, msig_m, [k[1] for k in ks], mult, m)
        assert not bgls.KoskVerifyMultiSignatureWithMultiplicity(curve, msig, [k[1] for k in ks], mult, m)
        assert not bgls.KoskVerifyMultiSignatureWithMultiplicity(curve, msig_m, [k[1] for k in ks], mult[:-1], m)
        dsigs = [bgls.DistinctMsgSign(curve, k[0], mm) for k, mm in zip(keys, msgs)]
        assert bgls.DistinctMsgVerifySingleSignature(curve, dsigs[0], pubs[0], msgs[0])
        dagg = bgls.AggregateSignatures(dsigs)
        assert bgls.DistinctMsgVerifyAggregateSignature(curve, dagg, pubs, msgs)
        assert not bgls.DistinctMsgVerifyAggregateSignature(curve, dagg, pubs, msgs[::-1])
        # many independent aggregate verifies in one batched launch: same verdicts as one call each
        items = [(agg, pubs, msgs), (agg2, pubs, msgs), (agg, pubs[:N - 1], msgs), (agg, pubs, [msgs[1], msgs[0]] + msgs[2:]),
                 (agg2, pubs + [vkf], msgs + [msgs[0]]), (dagg, pubs, msgs)]
        assert bgls.VerifyAggregateSignatures(curve, items) == [bgls.VerifyAggregateSignature(curve, *it) for it in items]
        assert bgls.VerifyAggregateSignatures(curve, items)[:2] == [True, False]
        # TestMul / TestAggregation of the curves package on the mirror (curves/curve_test.go:120-186)
        k = rng.randrange(curve.GetG1Order())
        inf, ok = curve.GetG1().Mul(k).Add(curve.GetG1().Mul(-k))
        assert ok and inf.Equals(curve.GetG1Infinity())
        gt, ok = curve.PairingProduct([curve.GetG1()], [curve.GetG2(), curve.GetG2()])
        assert gt is None and not ok      # length mismatch (curves/curve.go:126-128)
        assert curve.Pair(curve.GetG2(), curve.GetG2()) == (None, False)   # type mismatch (curves/altbn128.go:131-140)
        assert curve.Pair(curve.GetG1(), curve.GetG2Infinity())[0].Equals(curve.GetGTIdentity())


@pytest.mark.parametrize("cid,c", CURVES)
def test_concurrent_calls_from_host_threads(ctx, cid, c):
    """PairingProduct is called from many goroutines in the reference (curves/curve.go:132-134): concurrent
    host-buffer calls on one context run on separate execution slots and must return the same bytes as the
    serial calls, for products of different sizes in flight at the same time."""
    import threading
    rng = random.Random(77 + cid)
    cases = []
    for n in (1, 3, 8, 17, 40, 64):
        g1, _ = rand_points(cid, c, 1, n, rng)
        g2, _ = rand_points(cid, c, 2, n, rng)
        cases.append((g1, g2, n, C.pairing_product(cid, g1, g2, n, 4, 0)))
    agg1, agg2 = make_aggregate(cid, c, 12, rng)
    errors = []

    def work(tid):
        try:
            for rep in range(6):
                g1, g2, n, want = cases[(tid + rep) % len(cases)]
                gt, ok = ctx.pairing_product(cid, g1, g2, n)
                assert gt == want and not ok
                gt, ok = ctx.pairing_product(cid, agg1, agg2, 13)
                assert ok
                assert ctx.aggregate_points(cid, 2, g2, n) == C.aggregate(cid, 2, g2, n, 2)
        except Exception as e:  # noqa: BLE001
            errors.append(repr(e))
    ths = [threading.Thread(target=work, args=(t,)) for t in range(6)]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    assert not errors, errors


@pytest.mark.parametrize("cid,c", CURVES)
def test_compressed_wire_formats(ctx, cid, c):
    """Point.Marshal / UnmarshalG1 / UnmarshalG2 on compressed input (curves/altbn128.go:81-89,203-221,296-376;
    bls12_381.go:242-264) on the GPU: golden vectors, oracle parity on random and invalid records, ragged batch."""
    import json
    import os
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "compressed_golden.json")))[c.name]
    nb = c.nbytes
    for group, key in ((1, "g1"), (2, "g2")):
        unc = b"".join(bytes.fromhex(e["uncompressed"]) for e in gold[key])
        cmp_ = b"".join(bytes.fromhex(e["compressed"]) for e in gold[key])
        n = len(gold[key])
        assert ctx.compress_points(cid, group, unc, n) == cmp_
        pts, ok = ctx.decompress_points(cid, group, cmp_, n, check_subgroup=True)
        assert pts == unc and all(ok)
    rng = random.Random(404 + cid)
    flag = 0x80 if cid == 1 else 0
    for group in (1, 2):
        recs, want_pts, want_ok = [], [], []
        dec = c.decompress_g1 if group == 1 else c.decompress_g2
        marshal = c.marshal_g1 if group == 1 else c.marshal_g2
        for i in range(75):   # random abscissas: about half are not on the curve
            d = bytearray(rng.randrange(c.p).to_bytes(nb, "big") if group == 1 else
                          rng.randrange(c.p).to_bytes(nb, "big") + rng.randrange(c.p).to_bytes(nb, "big"))
            d[0] |= flag | (0x20 if (cid == 1 and i & 1) else 0) | (0x80 if (cid == 0 and i & 1 and d[0] < 0x30) else 0)
            pt, ok = dec(bytes(d))
            recs.append(bytes(d))
            want_ok.append(ok)
            want_pts.append(marshal(pt) if ok else bytes(2 * group * nb))
        pts, ok = ctx.decompress_points(cid, group, b"".join(recs), len(recs))
        assert ok == want_ok and pts == b"".join(want_pts)
        assert 10 < sum(ok) < 65
    # the curve mirror: Marshal / Unmarshal round trip, wrong length rejected (curves/curve_test.go:23-87)
    from bgls_b200.curves import Altbn128, Bls12
    curve = Altbn128 if cid == 0 else Bls12
    for gen, un in ((curve.GetG1(), curve.UnmarshalG1), (curve.GetG2(), curve.UnmarshalG2)):
        p = gen.Mul(rng.randrange(1, c.r))
        q, ok = un(p.Marshal())
        assert ok and q.Equals(p) and len(p.Marshal()) * 2 == len(p.MarshalUncompressed())
        q, ok = un(p.MarshalUncompressed())
        assert ok and q.Equals(p)
        assert un(p.Marshal()[:-1]) == (None, False)


@pytest.mark.parametrize("cid,c", CURVES)
def test_verify_aggregate_signature_entry_point(ctx, cid, c):
    """bgls_verify_aggregate_signature = verifyAggSig (bgls/bgls.go:94-119) in one call: accept / reject cases of
    bgls/bgls_test.go:40-77 with real HashToG1 points from the oracle, the n = 0 quirk, an infinity signature,
    ragged and empty messages, and the duplicate-message rule."""
    rng = random.Random(555 + cid)
    n = 9
    msgs = [bytes(rng.randrange(256) for _ in range(ln)) for ln in (32, 0, 1, 64, 31, 33, 100, 7, 32)]
    sks = [rng.randrange(1, c.r) for _ in range(n)]
    hs = [c.hash_to_g1(m) for m in msgs]
    pks = [c.g2_mul(c.g2, k) for k in sks]
    sigma = None
    for h, k in zip(hs, sks):
        sigma = c.g1_add(sigma, c.g1_mul(h, k))
    keys = b"".join(c.marshal_g2(q) for q in pks)
    sig = c.marshal_g1(sigma)
    V = ctx.verify_aggregate_signature
    assert V(cid, msgs, keys, sig) is True
    assert V(cid, msgs[::-1], keys, sig) is False                                   # swapped messages
    assert V(cid, msgs[:-1], keys[:-4 * c.nbytes], sig) is False                    # a signer dropped
    assert V(cid, msgs, keys, c.marshal_g1(c.g1_add(sigma, c.g1))) is False         # wrong signature
    # duplicate message: rejected before any pairing unless duplicates are allowed (Kosk / DistinctMsg callers)
    m2 = msgs[:-1] + [msgs[0]]
    sig2 = None
    for m, k in zip(m2, sks):
        sig2 = c.g1_add(sig2, c.g1_mul(c.hash_to_g1(m), k))
    assert V(cid, m2, keys, c.marshal_g1(sig2)) is False
    assert V(cid, m2, keys, c.marshal_g1(sig2), allow_duplicates=True) is True
    # n = 0: the product is the single pair (-sigma, g2): true iff sigma is the point at infinity (bgls.go:103-114)
    assert V(cid, [], b"", c.marshal_g1(None)) is True
    assert V(cid, [], b"", c.marshal_g1(c.g1)) is False
    # single signer
    assert V(cid, msgs[:1], keys[:4 * c.nbytes], c.marshal_g1(c.g1_mul(hs[0], sks[0]))) is True


@pytest.mark.parametrize("cid,c", CURVES)
def test_verify_multi_signature_entry_point(ctx, cid, c):
    """bgls_verify_multi_signature = verifyMultiSignature (bgls/bgls.go:89-92): bgls/blsKosk_test.go:35-64 shape."""
    rng = random.Random(808 + cid)
    msg = b"\x01" + bytes(rng.randrange(256) for _ in range(64))
    h = c.hash_to_g1(msg)
    for n in (1, 2, 8, 37):
        sks = [rng.randrange(1, c.r) for _ in range(n)]
        keys = b"".join(c.marshal_g2(c.g2_mul(c.g2, k)) for k in sks)
        sig = c.marshal_g1(c.g1_mul(h, sum(sks) % c.r))
        assert ctx.verify_multi_signature(cid, msg, keys, n, sig) is True
        assert ctx.verify_multi_signature(cid, msg + b"x", keys, n, sig) is False
        if n > 1:
            assert ctx.verify_multi_signature(cid, msg, keys[4 * c.nbytes:], n - 1, sig) is False
        assert ctx.verify_multi_signature(cid, msg, keys, n, c.marshal_g1(c.g1_add(c.unmarshal_g1(sig), c.g1))) is False


@pytest.mark.parametrize("cid,c", CURVES)
@pytest.mark.parametrize("engine", ["auto", "slot"])
def test_peer_memory_exchange_single_rank(cid, c, engine):
    """The peer-memory exchange path (bgls_exchange_*, bgls_miller_product_exchange_dev, bgls_final_exp_exchanged_dev)
    with world = 1: mailbox, send, flag wait and finish, several epochs on two lanes (both parities), against
    PairingProduct on the same inputs.  engine = auto: separate send / wait / guard kernels around the machine;
    engine = slot: the stores into the mailboxes close the Miller launch and the wait opens the finishing launch."""
    import os

    import torch

    import bgls_b200
    old = os.environ.get("BGLS_ENGINE")
    if engine == "slot":
        os.environ["BGLS_ENGINE"] = "slot"
    try:
        ctx2 = bgls_b200.Context(0)
    finally:
        if engine == "slot":
            if old is None:
                del os.environ["BGLS_ENGINE"]
            else:
                os.environ["BGLS_ENGINE"] = old
    try:
        h = ctx2.exchange_create(1, 0, 2)
        ctx2.exchange_connect(0, h)
        rng = random.Random(12 + cid)
        dev = torch.device("cuda", 0)
        s = torch.cuda.current_stream().cuda_stream
        F = c.nbytes
        d_out = torch.zeros(12 * F, dtype=torch.uint8, device=dev)
        d_flag = torch.zeros(1, dtype=torch.int32, device=dev)
        for epoch in (1, 2, 3):
            for lane, n in ((0, 5), (1, 40)):
                if epoch == 2:
                    g1, g2 = make_aggregate(cid, c, n - 1, rng)
                else:
                    g1, _ = rand_points(cid, c, 1, n, rng)
                    g2, _ = rand_points(cid, c, 2, n, rng)
                want, want_ok = ctx2.pairing_product(cid, g1, g2, n)
                t1 = torch.frombuffer(bytearray(g1), dtype=torch.uint8).to(dev)
                t2 = torch.frombuffer(bytearray(g2), dtype=torch.uint8).to(dev)
                ctx2.miller_product_exchange_dev(cid, t1.data_ptr(), t2.data_ptr(), n, lane, epoch, s)
                ctx2.final_exp_exchanged_dev(cid, lane, epoch, d_out.data_ptr(), d_flag.data_ptr(), s)
                torch.cuda.synchronize()
                assert bytes(d_out.cpu().numpy()) == want and bool(d_flag.item()) == want_ok
                assert want_ok == (epoch == 2)
        assert ctx2.exchange_error() == 0
    finally:
        ctx2.close()


@pytest.mark.parametrize("cid,c", CURVES)
def test_validate_points_rejects_what_the_reference_rejects(ctx, cid, c):
    """MakeG1Point / MakeG2Point / Unmarshal semantics (curves/altbn128.go:42-57,160-179; bls12_381.go:197-226,242-264):
    off-curve, unreduced (>= q) and out-of-subgroup records are refused before a Point exists."""
    from bgls_b200.curves import Altbn128, Bls12
    crv = Altbn128 if cid == 0 else Bls12
    nb, p = c.nbytes, c.p
    be = lambda v: int(v).to_bytes(nb, "big")
    rng = random.Random(90 + cid)
    good1, _ = rand_points(cid, c, 1, 3, rng)
    good2, _ = rand_points(cid, c, 2, 3, rng)
    # G1: valid points, infinity, off-curve (y + 1), unreduced x + q (does not fit for bls12: x + q < 2^384 always fits)
    P = c.unmarshal_g1(good1[:2 * nb])
    recs1 = [good1[:2 * nb], bytes(2 * nb), be(P[0]) + be((P[1] + 1) % p), be(P[0] + p if (P[0] + p).bit_length() <= 8 * nb else p) + be(P[1])]
    exp1 = [True, True, False, False]
    # a point on the curve outside the order-r subgroup (bls12 G1 has a cofactor; altbn128 G1 does not)
    if cid == 1:
        x = 5
        while True:
            y2 = (x ** 3 + c.b) % p
            y = pow(y2, (p + 1) // 4, p)
            if y * y % p == y2 and c.g1_mul((x, y), c.r) is not None:
                break
            x += 1
        recs1.append(be(x) + be(y))
        exp1.append(False)
    assert ctx.validate_points(cid, 1, b"".join(recs1), len(recs1)) == exp1
    assert ctx.validate_points(cid, 1, b"".join(recs1), len(recs1), reference=False) == exp1[:4] + [True] * (len(exp1) - 4)
    # G2: valid, infinity, off-curve, and a twist point outside the subgroup (both curves have a G2 cofactor)
    Q = c.unmarshal_g2(good2[:4 * nb])
    bad_y = c.marshal_g2((Q[0], ((Q[1][0] + 1) % p, Q[1][1])))
    F2 = c.F2
    xr = (3, 1)
    while True:
        y2 = F2.add(F2.mul(xr, F2.mul(xr, xr)), c.b2)
        yr = c.sqrt_fp2(y2)
        if F2.mul(yr, yr) == y2 and c.g2_mul((xr, yr), c.r) is not None:
            break
        xr = (xr[0] + 1, xr[1])
    recs2 = [good2[:4 * nb], bytes(4 * nb), bad_y, c.marshal_g2((xr, yr))]
    assert ctx.validate_points(cid, 2, b"".join(recs2), 4) == [True, True, False, False]
    assert ctx.validate_points(cid, 2, b"".join(recs2), 4, reference=False) == [True, True, False, True]
    # the host mirror applies them where the reference does
    pt, ok = crv.MakeG1Point([P[0], (P[1] + 1) % p], check=True)
    assert pt is None and not ok
    pt, ok = crv.MakeG1Point([P[0], P[1]], check=True)
    assert ok and pt.raw == good1[:2 * nb]
    pt, ok = crv.UnmarshalG2(recs2[3])
    assert pt is None and not ok
    pt, ok = crv.UnmarshalG2(recs2[2])
    assert pt is None and not ok
    pt, ok = crv.UnmarshalG1(recs1[2])
    assert pt is None and not ok
    pt, ok = crv.UnmarshalG2(good2[:4 * nb])
    assert ok
    if cid == 1:   # bls12-381 MakeG*Point(check = false) skips Check() (bls12_381.go:203,222)
        pt, ok = crv.MakeG1Point([P[0], (P[1] + 1) % p], check=False)
        assert ok


@pytest.mark.parametrize("cid,c", CURVES)
def test_gt_pow_is_point_t_mul(ctx, cid, c):
    """PointT.Mul (curves/curve.go:63-70): e(P, Q)^k == e(kP, Q); k = 0, 1, -1, r-1 and random."""
    from bgls_b200.curves import Altbn128, Bls12
    crv = Altbn128 if cid == 0 else Bls12
    rng = random.Random(17 + cid)
    g1, ks = rand_points(cid, c, 1, 1, rng)
    g2, _ = rand_points(cid, c, 2, 1, rng)
    base = ctx.pair(cid, c.marshal_g1(c.g1), g2)
    assert ctx.gt_pow(cid, base, ks[0]) == ctx.pair(cid, g1, g2)
    one = c.marshal_gt(c.fp12_one)
    assert ctx.gt_pow(cid, base, 0) == one and ctx.gt_pow(cid, base, 1) == base
    inv = ctx.gt_pow(cid, base, -1)
    assert ctx.gt_mul(cid, inv, base) == one and ctx.gt_pow(cid, base, c.r - 1) == inv
    t = crv.UnmarshalGT(base)[0]
    assert t.Mul(ks[0]).Marshal() == ctx.pair(cid, g1, g2)
