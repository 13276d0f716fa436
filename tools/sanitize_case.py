"""Mid-size cases for compute-sanitizer runs (fused in-block product tree with several warps per block, batch path):
    compute-sanitizer --tool racecheck python tools/sanitize_case.py"""
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bgls_b200  # noqa: E402
from parity_util import CURVES, make_aggregate  # noqa: E402

ctx = bgls_b200.Context(0)
for cid, c in CURVES:
    g1, g2 = make_aggregate(cid, c, 449, random.Random(7 + cid), nthreads=8)   # 450 pairs: 4 warps per block
    gt, ok = ctx.pairing_product(cid, g1, g2, 450)
    assert ok and gt == c.marshal_gt(c.fp12_one)
    nb = c.nbytes
    offs = [0, 150, 300, 450]
    res = ctx.pairing_check_batch(cid, g1, g2, offs)
    assert res == [False, False, False], res   # parts of a valid aggregate are not valid on their own
    assert len(ctx.aggregate_points(cid, 2, g2, 450)) == 4 * nb
print("sanitize cases ok")
