#!/usr/bin/env python
"""Where the time of BASELINE config 3 goes (bls12-381 verify_multi_signature, 65,536 G2 keys from pinned host
memory): the whole call, and its parts timed alone -- the hash of the one message, the key aggregation, the
host-to-device copy, the 2-pair check.  python tools/bench_c3.py"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import numpy as np

import bgls_b200
from bench import ORDER, SEED
from bgls_b200.curves import Altbn128, Bls12

ctx = bgls_b200.Context(0)
dev = torch.device("cuda", 0)
out = {}
for cid, name, F in ((1, "bls12-381", 48), (0, "altbn128", 32)):
    n, BASE = 65536, 1024
    rs = np.random.RandomState((SEED + 3000) & 0x7FFFFFFF)
    raw = rs.bytes(32 * BASE)
    ss = [int.from_bytes(raw[32 * i:32 * i + 32], "big") % ORDER[cid] for i in range(BASE)]
    crv = Altbn128 if cid == 0 else Bls12
    pk = ctx.scale_points(cid, 2, crv.GetG2().raw * BASE, b"".join(k.to_bytes(32, "big") for k in ss), BASE)
    keys = pk * (n // BASE)
    msg = b"\x01" + b"config 3: one message, 65536 signers"
    H = ctx.hash_to_g1(cid, [msg])
    sig = ctx.scale_points(cid, 1, H, (((n // BASE) * sum(ss)) % ORDER[cid]).to_bytes(32, "big"), 1)
    h_keys = torch.frombuffer(bytearray(keys), dtype=torch.uint8).pin_memory()

    def wall(fn, reps=8):
        ts = []
        for _ in range(reps):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            fn()
            ts.append((time.perf_counter() - t0) * 1e3)
        return {"min_ms": min(ts[1:]), "median_ms": sorted(ts[1:])[len(ts[1:]) // 2]}

    r = {}
    r["verify_multi_signature"] = wall(lambda: ctx.verify_multi_signature_ptr(cid, msg, h_keys.data_ptr(), n, sig))
    r["hash_one_message"] = wall(lambda: ctx.hash_to_g1(cid, [msg]))
    d_keys = torch.empty_like(h_keys, device=dev)
    r["h2d_keys"] = wall(lambda: (d_keys.copy_(h_keys, non_blocking=True), torch.cuda.synchronize()))
    d_sum = torch.zeros(4 * F, dtype=torch.uint8, device=dev)
    s = torch.cuda.current_stream().cuda_stream
    r["aggregate_dev"] = wall(lambda: (ctx.aggregate_points_dev(cid, 2, d_keys.data_ptr(), n, d_sum.data_ptr(), s), torch.cuda.synchronize()))
    agg = bytes(d_sum.cpu().numpy().tobytes())
    g2 = ctx.aggregate_points(cid, 2, keys[:4 * F], 1)   # any second G2 point: the time of a 2-pair product does not depend on the verdict
    r["pairing_check_2_pairs"] = wall(lambda: ctx.pairing_product(cid, H + sig, agg + g2, 2))
    out[name] = r
print(json.dumps(out))
