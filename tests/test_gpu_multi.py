"""Two GPUs, one process each (NCCL): the sharded aggregate verify of SURVEY.md 8e -- per-rank Miller product, one
all-gather of the Fp12 partials, final exponentiation on every rank -- against the oracle: GT bytes and verdict of a valid
and of a corrupted aggregate.  Skipped on a box with fewer than two GPUs."""
import os
import random
import socket
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, cid, g1, g2, n, want, q):
    try:
        import torch
        import torch.distributed as dist
        sys.path.insert(0, ROOT)
        import bgls_b200
        from bgls_b200.sharded import shard_bounds, sharded_pairing_product
        os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
        torch.cuda.set_device(rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
        ctx = bgls_b200.Context(rank)
        F = 32 if cid == 0 else 48
        lo, hi = shard_bounds(n, world, rank)
        res = []
        for a, b in zip(g1, g2):
            gt, ok = sharded_pairing_product(ctx, cid, a[2 * F * lo:2 * F * hi], b[4 * F * lo:4 * F * hi], hi - lo,
                                             device=torch.device("cuda", rank))
            res.append((gt, bool(ok)))
        ctx.close()
        dist.destroy_process_group()
        q.put((rank, res == want, None))
    except Exception as e:  # noqa: BLE001
        q.put((rank, False, repr(e)))


@pytest.mark.parametrize("cid", [0, 1])
def test_two_rank_sharded_verify(cid):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp

    from oracle import c_oracle as C
    from parity_util import CURVES, make_aggregate
    c = dict(CURVES)[cid]
    rng = random.Random(404 + cid)
    n = 41
    g1, g2 = make_aggregate(cid, c, n - 1, rng)
    nb = c.nbytes
    bad1 = g1[2 * nb:4 * nb] + g1[:2 * nb] + g1[4 * nb:]
    want = [(C.pairing_product(cid, a, b, n, 8, 0),) for a, b in ((g1, g2), (bad1, g2))]
    one = c.marshal_gt(c.fp12_one)
    want = [(w[0], w[0] == one) for w in want]
    assert want[0][1] and not want[1][1]
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mpc = mp.get_context("spawn")
    q = mpc.Queue()
    ps = [mpc.Process(target=_worker, args=(r, 2, port, cid, [g1, bad1], [g2, g2], n, want, q)) for r in range(2)]
    for p in ps:
        p.start()
    out = [q.get(timeout=600) for _ in ps]
    for p in ps:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in out), out
