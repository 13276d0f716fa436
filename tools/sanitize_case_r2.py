"""Round-2 kernels under compute-sanitizer: the slot engine (in-block tree, cross-block ticket tree, final exponentiation
in the Miller launch, batch form), the six-lane aggregation (block tree + cross-block tree), the pooled hash-to-G1,
validation and GT exponentiation.
    BGLS_ENGINE=slot BGLS_HASH=pool compute-sanitizer --tool racecheck python tools/sanitize_case_r2.py"""
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ.setdefault("BGLS_ENGINE", "slot")
os.environ.setdefault("BGLS_HASH", "pool")
import bgls_b200  # noqa: E402
from oracle import c_oracle as C  # noqa: E402
from parity_util import CURVES, make_aggregate, rand_points  # noqa: E402

ctx = bgls_b200.Context(0)
for cid, c in CURVES:
    n = 300                                                                   # 38 one-warp blocks: two levels of the cross-block tree
    g1, g2 = make_aggregate(cid, c, n - 1, random.Random(17 + cid), nthreads=8)
    gt, ok = ctx.pairing_product(cid, g1, g2, n)
    assert ok and gt == c.marshal_gt(c.fp12_one)
    res = ctx.pairing_check_batch(cid, g1, g2, [0, 100, 100, 101, n])         # ragged batch incl. an empty and a one-pair product
    assert res == [False, True, False, False], res
    for group in (1, 2):
        m = 900                                                               # 45 blocks of 20 groups: 45 -> 3 -> 1
        pts, _ = rand_points(cid, c, group, m, random.Random(3 + group), nthreads=8)
        assert ctx.aggregate_points(cid, group, pts, m) == C.aggregate(cid, group, pts, m, 8)
    assert all(ctx.validate_points(cid, 2, g2[:4 * c.nbytes * 8], 8, True))
    assert ctx.gt_pow(cid, c.marshal_gt(c.fp12_one), 5) == c.marshal_gt(c.fp12_one)
msgs = [bytes([i & 255, i >> 8]) * (1 + i % 7) for i in range(70)]
h = ctx.hash_to_g1(0, msgs)
c0 = CURVES[0][1]
assert h[:64] == c0.marshal_g1(c0.hash_to_g1(msgs[0])) and h[69 * 64:] == c0.marshal_g1(c0.hash_to_g1(msgs[69]))
print("sanitize cases (round 2) ok")
