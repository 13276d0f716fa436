#!/usr/bin/env python
"""BASELINE config 5 on N GPUs: 4096 independent altbn128 aggregate verifies x 256 signers (257 pairs each), dealt to
the ranks in contiguous runs (no data-path collective), the verdict mask all-gathered over NCCL.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tools/bench_multi.py [--checks 4096] [--signers 256] [--reps 5]

Device-resident inputs, CUDA events on the launching stream, barrier + synchronize on both sides, max over ranks.
(Config 4 -- one 2^20-signer bls12-381 product sharded over 8 GPUs -- is bench.py --gpus 8 --curve bls12 --signers 131072.)"""
import argparse
import json
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--checks", type=int, default=4096)
    ap.add_argument("--signers", type=int, default=256)
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist

    import bgls_b200
    from bgls_b200.curves import Altbn128
    from bgls_b200.sharded import shard_bounds
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG", "WARN")
        dist.init_process_group("nccl", device_id=dev)
    ctx = bgls_b200.Context(local)
    cid, curve, r = 0, Altbn128, 21888242871839275222246405745257275088548364400416034343698204186575808495617
    rng = random.Random(5)     # same stream on every rank: one logical batch
    to_b = lambda ks: b"".join(int(k).to_bytes(32, "big") for k in ks)
    S, B = args.signers, args.checks
    hs, ss = [rng.randrange(1, r) for _ in range(S)], [rng.randrange(1, r) for _ in range(S)]
    g1 = ctx.scale_points(cid, 1, curve.GetG1().raw * S, to_b(hs), S)
    g2 = ctx.scale_points(cid, 2, curve.GetG2().raw * S, to_b(ss), S)
    tot = sum(h * x for h, x in zip(hs, ss)) % r
    neg = ctx.scale_points(cid, 1, curve.GetG1().raw, to_b([(r - tot) % r]), 1)
    a1, a2, bad1 = g1 + neg, g2 + curve.GetG2().raw, g1 + curve.GetG1().raw
    expect = [(b % 7) != 3 for b in range(B)]
    lo, hi = shard_bounds(B, world, rank)
    nb = hi - lo
    G1 = torch.frombuffer(bytearray(b"".join(a1 if e else bad1 for e in expect[lo:hi])), dtype=torch.uint8).to(dev)
    G2 = torch.frombuffer(bytearray(a2 * nb), dtype=torch.uint8).to(dev)
    off = torch.arange(0, (nb + 1) * (S + 1), S + 1, dtype=torch.int64, device=dev)
    width = -(-B // world)
    ok = torch.full((width,), 255, dtype=torch.uint8, device=dev)
    mask = torch.zeros(world * width, dtype=torch.uint8, device=dev)
    s = torch.cuda.current_stream().cuda_stream

    def step():
        ctx.pairing_check_batch_dev(cid, G1.data_ptr(), G2.data_ptr(), off.data_ptr(), nb, nb * (S + 1), ok.data_ptr(), s)
        if world > 1:
            dist.all_gather_into_tensor(mask, ok)
        else:
            mask.copy_(ok)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    step()
    barrier()
    best = 1e30
    for _ in range(args.reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        step()
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        best = min(best, float(t.item()))
    got = []
    m = mask.cpu().tolist()
    for q in range(world):
        a, b = shard_bounds(B, world, q)
        got += [bool(x) for x in m[q * width:q * width + (b - a)]]
    assert got == expect, "batch verdicts differ"
    if rank == 0:
        total = B * (S + 1)
        print(json.dumps({"config5_batch": {"n_gpus": world, "checks": B, "pairs_per_check": S + 1, "ms": best,
                                            "pairings_per_s": total / (best * 1e-3), "checks_per_s": B / (best * 1e-3),
                                            "sharding": "contiguous runs of checks per rank, verdict mask all-gathered (NCCL)"}}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
