#!/usr/bin/env python
"""Randomised differential run of the C ABI against the C oracle (test infrastructure): products on every engine with
points at infinity and repeated / cancelling pairs sprinkled in, sharded products finished from k partials, aggregation
with infinities, doublings and cancelling points at sizes that straddle the block and tree boundaries, hash-to-G1 in both
forms.      python tools/fuzz_gpu.py [--seconds 120] [--seed 1]"""
import argparse
import os
import random
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bgls_b200  # noqa: E402
from oracle import c_oracle as C  # noqa: E402
from parity_util import CURVES, make_aggregate, rand_points, scalars_bytes  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--seconds", type=float, default=120.0)
ap.add_argument("--seed", type=int, default=1)
a = ap.parse_args()
rng = random.Random(a.seed)


def ctx_with(**env):
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        return bgls_b200.Context(0)
    finally:
        for k, v in old.items():
            if v is None:
                del os.environ[k]
            else:
                os.environ[k] = v


ctxs = {"auto": bgls_b200.Context(0), "slot": ctx_with(BGLS_ENGINE="slot", BGLS_HASH="pool"), "machine": ctx_with(BGLS_ENGINE="machine", BGLS_HASH="wide")}
t_end = time.time() + a.seconds
stats = {"products": 0, "sharded": 0, "aggregates": 0, "hashes": 0}
while time.time() < t_end:
    cid, c = rng.choice(CURVES)
    nb = c.nbytes
    # ---- a product with special pairs
    n = rng.choice([1, 2, 3, 7, 8, 9, 15, 16, 17, 31, 32, 33, 63, 64, 65, 100, 129, 257, 300])
    if rng.random() < 0.5 and n > 1:
        g1, g2 = make_aggregate(cid, c, n - 1, rng, nthreads=8)
    else:
        g1, _ = rand_points(cid, c, 1, n, rng)
        g2, _ = rand_points(cid, c, 2, n, rng)
    g1, g2 = bytearray(g1), bytearray(g2)
    for _ in range(rng.choice([0, 0, 1, 2, 5])):
        i = rng.randrange(n)
        kind = rng.randrange(4)
        if kind == 0:
            g1[2 * nb * i:2 * nb * (i + 1)] = bytes(2 * nb)                      # G1 at infinity
        elif kind == 1:
            g2[4 * nb * i:4 * nb * (i + 1)] = bytes(4 * nb)                      # G2 at infinity
        elif kind == 2:
            j = rng.randrange(n)
            g1[2 * nb * i:2 * nb * (i + 1)] = g1[2 * nb * j:2 * nb * (j + 1)]    # repeated G1 point
        else:
            j = rng.randrange(n)
            g2[4 * nb * i:4 * nb * (i + 1)] = g2[4 * nb * j:4 * nb * (j + 1)]    # repeated G2 point
    g1, g2 = bytes(g1), bytes(g2)
    want = C.pairing_product(cid, g1, g2, n, 8, 0)
    for name, ctx in ctxs.items():
        gt, one = ctx.pairing_product(cid, g1, g2, n)
        assert gt == want and one == (want == c.marshal_gt(c.fp12_one)), ("product", name, c.name, n)
    stats["products"] += 1
    # ---- the same product from k shards, finished by every engine
    k = rng.randrange(1, min(n, 9) + 1)
    cuts = sorted(rng.sample(range(1, n), k - 1)) if k > 1 else []
    cuts = [0] + cuts + [n]
    for name, ctx in ctxs.items():
        parts = b"".join(ctx.miller_product(cid, g1[2 * nb * lo:2 * nb * hi], g2[4 * nb * lo:4 * nb * hi], hi - lo) for lo, hi in zip(cuts, cuts[1:]))
        gt, one = ctx.final_exp_product(cid, parts, k)
        assert gt == want, ("sharded", name, c.name, n, k)
    stats["sharded"] += 1
    # ---- aggregation with special points
    group = rng.choice([1, 2])
    rec = 2 * group * nb
    m = rng.choice([1, 2, 3, 5, 19, 20, 21, 39, 40, 41, 100, 399, 400, 401, 1000, 3000])
    pts, ks = rand_points(cid, c, group, m, rng)
    pts = bytearray(pts)
    gen = c.marshal_g1(c.g1) if group == 1 else c.marshal_g2(c.g2)
    for _ in range(rng.choice([0, 1, 3, 10])):
        i, j = rng.randrange(m), rng.randrange(m)
        kind = rng.randrange(3)
        if kind == 0:
            pts[rec * i:rec * (i + 1)] = bytes(rec)
        elif kind == 1:
            pts[rec * i:rec * (i + 1)] = pts[rec * j:rec * (j + 1)]
        else:
            pts[rec * i:rec * (i + 1)] = C.scale_points(cid, group, gen, scalars_bytes([(c.r - ks[j]) % c.r]), 1)
    pts = bytes(pts)
    want = C.aggregate(cid, group, pts, m, 8)
    assert ctxs["auto"].aggregate_points(cid, group, pts, m) == want, ("aggregate", c.name, group, m)
    stats["aggregates"] += 1
    # ---- hash-to-G1, both forms
    msgs = [bytes(rng.randrange(256) for _ in range(rng.choice([0, 1, 31, 32, 33, 64, 100, 136, 137]))) for _ in range(rng.choice([1, 3, 33, 70]))]
    h = ctxs["slot"].hash_to_g1(cid, msgs)
    assert h == ctxs["machine"].hash_to_g1(cid, msgs), ("hash forms", c.name)
    i = rng.randrange(len(msgs))
    assert h[2 * nb * i:2 * nb * (i + 1)] == c.marshal_g1(c.hash_to_g1(msgs[i])), ("hash", c.name)
    stats["hashes"] += len(msgs)
for ctx in ctxs.values():
    ctx.close()
print("fuzz ok", stats)
