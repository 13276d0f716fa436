#!/usr/bin/env python
"""Times the hash-to-G1 kernel (device-resident messages, CUDA events): python tools/bench_hash.py"""
import json
import os
import random
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bgls_b200

ctx = bgls_b200.Context(0)
dev = torch.device("cuda", 0)
rng = np.random.RandomState(1)
out = {}
for cid, name, F in ((0, "altbn128", 32), (1, "bls12", 48)):
    for n in (1024, 16384, 131072):
        msgs = torch.from_numpy(rng.randint(0, 256, size=32 * n, dtype=np.uint8)).to(dev)
        offs = torch.arange(0, 32 * (n + 1), 32, dtype=torch.int64, device=dev)
        pts = torch.zeros(n * 2 * F, dtype=torch.uint8, device=dev)
        s = torch.cuda.current_stream().cuda_stream
        best = 1e9
        for rep in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            rc = ctx._L.bgls_hash_to_g1_dev(ctx._h, cid, msgs.data_ptr(), offs.data_ptr(), n, pts.data_ptr(), s)
            assert rc == 0
            e1.record()
            torch.cuda.synchronize()
            if rep:
                best = min(best, e0.elapsed_time(e1))
        out["%s_%d" % (name, n)] = {"ms": best, "hashes_per_s": n / (best * 1e-3)}
print(json.dumps(out))
