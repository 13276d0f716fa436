"""N > 1 orchestration on CPU: world_size-2 gloo, with the C oracle plugged in as the byte-level engine
(test infrastructure standing in for the CUDA context) -- checks the slicing, the all-gather of the
Fp12 partials / partial sums and the final combination against the unsharded oracle result."""
import os
import random
import sys

import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class OracleEngine:
    def miller_product(self, curve, g1, g2, n):
        from oracle import c_oracle as C
        return C.miller_product(curve, g1, g2, n, 1)

    def final_exp_product(self, curve, parts, k):
        from oracle import c_oracle as C
        from oracle import bgls_oracle as O
        gt = C.fp12_product(curve, parts, k, True)
        c = O.ALTBN128 if curve == 0 else O.BLS12_381
        return gt, gt == c.marshal_gt(c.fp12_one)

    def aggregate_points(self, curve, grp, pts, n):
        from oracle import c_oracle as C
        return C.aggregate(curve, grp, pts, n, 1)

    def pairing_check_batch(self, curve, g1, g2, offsets):
        from oracle import c_oracle as C
        from oracle import bgls_oracle as O
        c = O.ALTBN128 if curve == 0 else O.BLS12_381
        nb, one = c.nbytes, c.marshal_gt(c.fp12_one)
        return [C.pairing_product(curve, g1[2 * nb * a:2 * nb * b], g2[4 * nb * a:4 * nb * b], b - a, 1, 0) == one
                for a, b in zip(offsets[:-1], offsets[1:])]


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    from bgls_b200 import sharded
    from oracle import c_oracle as C
    from parity_util import CURVES, make_aggregate, rand_points
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    res = {}
    try:
        for cid, c in CURVES:
            rng = random.Random(99 + cid)  # same stream on every rank: identical global arrays
            n = 9
            g1, g2 = make_aggregate(cid, c, n, rng, nthreads=1)
            lo, hi = sharded.shard_bounds(n + 1, world, rank)
            nb = c.nbytes
            gt, ok = sharded.sharded_pairing_product(OracleEngine(), cid, g1[2 * nb * lo:2 * nb * hi], g2[4 * nb * lo:4 * nb * hi], hi - lo)
            res[f"pair{cid}"] = (ok, gt == C.pairing_product(cid, g1, g2, n + 1, 1, 0))
            pts, _ = rand_points(cid, c, 2, 7, rng, nthreads=1)
            lo, hi = sharded.shard_bounds(7, world, rank)
            rec = 4 * nb
            s = sharded.sharded_aggregate_points(OracleEngine(), cid, 2, pts[rec * lo:rec * hi], hi - lo)
            res[f"agg{cid}"] = s == C.aggregate(cid, 2, pts, 7, 1)
            # batch of 5 ragged checks (sizes 3, 1, 0, 2, 4 pairs), checks 1 and 3 corrupted: dealt 2 + 3 over the ranks
            G1, G2, offs, expect = [], [], [0], []
            for b, m in enumerate((2, 0, None, 1, 3)):
                if m is None:          # an empty product is the identity
                    offs.append(offs[-1])
                    expect.append(True)
                    continue
                a1, a2 = make_aggregate(cid, c, m, rng, nthreads=1)
                if b in (1, 3):
                    a1 = a1[:-2 * nb] + c.marshal_g1(c.g1)
                G1.append(a1)
                G2.append(a2)
                offs.append(offs[-1] + m + 1)
                expect.append(b not in (1, 3))
            got = sharded.sharded_pairing_check_batch(OracleEngine(), cid, b"".join(G1), b"".join(G2), offs)
            res[f"batch{cid}"] = got == expect
    finally:
        dist.destroy_process_group()
    q.put((rank, res))


def test_sharded_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 200
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, res in out:
        for cid in (0, 1):
            assert res[f"pair{cid}"] == (True, True), (rank, cid, res)
            assert res[f"agg{cid}"], (rank, cid)
            assert res[f"batch{cid}"], (rank, cid)


def test_shard_bounds_cover():
    from bgls_b200.sharded import shard_bounds
    for n in (0, 1, 7, 1025):
        for w in (1, 2, 3, 8):
            cuts = [shard_bounds(n, w, r) for r in range(w)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(w - 1))
