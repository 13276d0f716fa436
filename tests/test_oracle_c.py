"""C oracle (oracle/liboracle.so) against the Python big-int oracle and the golden vectors."""
import json
import os
import random

import pytest

from oracle import bgls_oracle as O
from oracle import c_oracle as C

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "pairing_golden.json")))
CURVES = [(C.ALTBN128, O.ALTBN128), (C.BLS12, O.BLS12_381)]


@pytest.mark.parametrize("cid,c", CURVES)
def test_golden_pairing_vectors(cid, c):
    g = GOLD[c.name]
    g1, g2, n = bytes.fromhex(g["g1"]), bytes.fromhex(g["g2"]), g["n"]
    assert C.pairing_product(cid, g1, g2, n, 1, 0).hex() == g["product_gt"]
    assert C.pairing_product(cid, g1, g2, n, 3, 1).hex() == g["product_gt"]  # reference structure
    assert C.pairing_product(cid, c.marshal_g1(c.g1), c.marshal_g2(c.g2), 1).hex() == g["gen_gt"]
    nb = c.nbytes
    assert C.pairing_product(cid, g1[:2 * nb], g2[:4 * nb], 1).hex() == g["pair0_gt"]
    assert C.aggregate(cid, 1, g1, n, 2).hex() == g["sum_g1"]
    assert C.aggregate(cid, 2, g2, n, 2).hex() == g["sum_g2"]


@pytest.mark.parametrize("cid,c", CURVES)
def test_final_exp_matches_definition(cid, c):
    g = GOLD[c.name]
    g1, g2, n = bytes.fromhex(g["g1"]), bytes.fromhex(g["g2"]), g["n"]
    mf = C.miller_product(cid, g1, g2, n, 2)
    assert C.fp12_pow(cid, mf, (c.p ** 12 - 1) // c.r).hex() == g["product_gt"]
    assert C.fp12_product(cid, mf, 1, True).hex() == g["product_gt"]


@pytest.mark.parametrize("cid,c", CURVES)
def test_sharded_product_equals_unsharded(cid, c):
    """multi-GPU plan: per-shard Miller products multiplied, one final exp."""
    g = GOLD[c.name]
    g1, g2, n = bytes.fromhex(g["g1"]), bytes.fromhex(g["g2"]), g["n"]
    nb = c.nbytes
    parts = b"".join(C.miller_product(cid, g1[2 * nb * i:2 * nb * (i + 1)], g2[4 * nb * i:4 * nb * (i + 1)], 1) for i in range(n))
    assert C.fp12_product(cid, parts, n, True).hex() == g["product_gt"]


@pytest.mark.parametrize("cid,c", CURVES)
def test_agg_signature_boolean(cid, c):
    a = GOLD[c.name]["agg"]
    hs = [bytes.fromhex(h) for h in a["hashes"]]
    pks = [bytes.fromhex(h) for h in a["pubkeys"]]
    sig = c.unmarshal_g1(bytes.fromhex(a["sig"]))
    neg = c.marshal_g1(c.g1_neg(sig))
    one = c.marshal_gt(c.fp12_one)
    g1 = b"".join(hs) + neg
    g2 = b"".join(pks) + bytes.fromhex(a["g2gen"])
    assert C.pairing_product(cid, g1, g2, 4, 2, 0) == one
    # swapped messages -> reject
    g1bad = hs[1] + hs[0] + hs[2] + neg
    assert C.pairing_product(cid, g1bad, g2, 4, 2, 0) != one


@pytest.mark.parametrize("cid,c", CURVES)
def test_group_ops_random(cid, c):
    rng = random.Random(11)
    for grp, add, mul, gen, mar in ((1, c.g1_add, c.g1_mul, c.g1, c.marshal_g1), (2, c.g2_add, c.g2_mul, c.g2, c.marshal_g2)):
        pts = [mul(gen, rng.randrange(c.r)) for _ in range(6)]
        blob = b"".join(mar(P) for P in pts)
        s = None
        for P in pts:
            s = add(s, P)
        assert C.aggregate(cid, grp, blob, 6, 3) == mar(s)
        sc = [0, 1, c.r - 1] + [rng.randrange(c.r) for _ in range(3)]
        out = C.scale_points(cid, grp, blob, b"".join(k.to_bytes(32, "big") for k in sc), 6, 2)
        assert out == b"".join(mar(mul(P, k)) for P, k in zip(pts, sc))
        # infinity, doubling and cancellation inside a sum
        blob2 = mar(pts[0]) + mar(None) + mar(pts[0]) + mar(pts[1]) + mar(ec_neg(c, grp, pts[1]))
        assert C.aggregate(cid, grp, blob2, 5, 1) == mar(add(pts[0], pts[0]))
        assert C.on_curve(cid, grp, mar(pts[2]))


def ec_neg(c, grp, P):
    return c.g1_neg(P) if grp == 1 else c.g2_neg(P)


@pytest.mark.parametrize("cid,c", CURVES)
def test_infinity_pairs(cid, c):
    one = c.marshal_gt(c.fp12_one)
    assert C.pairing_product(cid, c.marshal_g1(None), c.marshal_g2(c.g2), 1) == one
    assert C.pairing_product(cid, c.marshal_g1(c.g1), c.marshal_g2(None), 1) == one
    assert C.pairing_product(cid, b"", b"", 0) == one
