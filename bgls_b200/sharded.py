"""Signer-set sharding across GPUs: one process per GPU, `torch.distributed` for the
plumbing.  Every rank multiplies the Miller values of its own contiguous slice of the packed
(H(m_i), pk_i) arrays; the per-rank Fp12 partials (12*F bytes each) are all-gathered -- the only
collective, pure latency over NVLink/NVSwitch -- and every rank finishes with one final
exponentiation, so all ranks hold the verdict.  Point aggregation shards the same way with
partial sums.  Independent checks (throughput mode, BASELINE config 5) are dealt to the ranks in
contiguous runs and need no data-path collective at all: only the verdict mask is gathered.

`engine` is anything with the byte-level methods of bgls_b200.Context (miller_product,
final_exp_product, aggregate_points): the CUDA context in production; the gloo CPU tests plug in
the oracle to exercise exactly this orchestration without a GPU.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from ._native import FP_BYTES


def _all_gather_bytes(blob: bytes, group=None, device=None) -> bytes:
    world = dist.get_world_size(group)
    t = torch.frombuffer(bytearray(blob), dtype=torch.uint8)
    if device is not None:
        t = t.to(device)
    out = torch.empty(world * len(blob), dtype=torch.uint8, device=t.device)
    dist.all_gather_into_tensor(out, t, group=group)
    return bytes(out.cpu().numpy().tobytes())


def shard_bounds(n: int, world: int, rank: int):
    """Contiguous slice [lo, hi) of n items owned by `rank`."""
    return n * rank // world, n * (rank + 1) // world


def sharded_pairing_product(engine, curve: int, g1_local: bytes, g2_local: bytes, n_local: int, group=None, device=None):
    """prod over all ranks' pairs of e(g1, g2): returns (gt_bytes, is_identity) on every rank."""
    partial = engine.miller_product(curve, g1_local, g2_local, n_local)
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return engine.final_exp_product(curve, partial, 1)
    parts = _all_gather_bytes(partial, group, device)
    return engine.final_exp_product(curve, parts, dist.get_world_size(group))


def sharded_aggregate_points(engine, curve: int, grp: int, pts_local: bytes, n_local: int, group=None, device=None):
    """sum over all ranks' points; a rank with an empty slice contributes the point at infinity."""
    rec = 2 * grp * FP_BYTES[curve]
    partial = engine.aggregate_points(curve, grp, pts_local, n_local) if n_local > 0 else bytes(rec)
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return partial
    parts = _all_gather_bytes(partial, group, device)
    return engine.aggregate_points(curve, grp, parts, dist.get_world_size(group))


def sharded_pairing_check_batch(engine, curve: int, g1: bytes, g2: bytes, offsets, group=None, device=None):
    """nbatch independent pairing-product checks (one verifyAggSig each, bgls/bgls.go:94-119) dealt to the ranks:
    rank r runs checks [lo, hi) of the batch through `engine.pairing_check_batch` on its own GPU -- no exchange on the
    data path -- and the verdict mask is all-gathered, so every rank returns the full list of nbatch booleans.
    `g1` / `g2` / `offsets` describe the WHOLE batch on every rank (pairs of check b are [offsets[b], offsets[b+1]))."""
    nb = len(offsets) - 1
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return engine.pairing_check_batch(curve, g1, g2, list(offsets)) if nb else []
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    F = FP_BYTES[curve]
    lo, hi = shard_bounds(nb, world, rank)
    local = []
    if hi > lo:
        p0, p1 = offsets[lo], offsets[hi]
        local = engine.pairing_check_batch(curve, g1[2 * F * p0:2 * F * p1], g2[4 * F * p0:4 * F * p1],
                                           [o - p0 for o in offsets[lo:hi + 1]])
    width = -(-nb // world) if nb else 0   # every rank contributes the same number of bytes: pad with 0xFF
    if width == 0:
        return []
    blob = bytes(int(b) for b in local) + b"\xff" * (width - len(local))
    parts = _all_gather_bytes(blob, group, device)
    out = []
    for r in range(world):
        a, b = shard_bounds(nb, world, r)
        out += [bool(x) for x in parts[r * width:r * width + (b - a)]]
    return out
