"""Short single-GPU command for ncu captures of the pairing kernels: a valid aggregate of --pairs pairs (synthesised on the
GPU as in bench.py) verified --reps times through bgls_pairing_product_dev.
    ncu --set full --clock-control none --import-source on -k regex:k_mach_miller32 -s 2 -c 1 -o gpurun_out/prof \
        python tools/ncu_case.py --pairs 1025            (BGLS_ENGINE=machine keeps the machine kernels at any size)"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bgls_b200  # noqa: E402
from bgls_b200.curves import Altbn128, Bls12  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--pairs", type=int, default=1025)
ap.add_argument("--curve", default="altbn128")
ap.add_argument("--reps", type=int, default=4)
a = ap.parse_args()
cid, curve = (0, Altbn128) if a.curve == "altbn128" else (1, Bls12)
r, S = curve.order, a.pairs - 1
ctx = bgls_b200.Context(0)
rs = np.random.RandomState(7)
raw = rs.bytes(64 * S)
hs = [int.from_bytes(raw[32 * i:32 * i + 32], "big") % r for i in range(S)]
ss = [int.from_bytes(raw[32 * (S + i):32 * (S + i) + 32], "big") % r for i in range(S)]
to_b = lambda ks: b"".join(k.to_bytes(32, "big") for k in ks)
g1 = ctx.scale_points(cid, 1, curve.GetG1().raw * S, to_b(hs), S)
g2 = ctx.scale_points(cid, 2, curve.GetG2().raw * S, to_b(ss), S)
tot = sum(h * s for h, s in zip(hs, ss)) % r
g1 += ctx.scale_points(cid, 1, curve.GetG1().raw, ((r - tot) % r).to_bytes(32, "big"), 1)
g2 += curve.GetG2().raw
dev = torch.device("cuda", 0)
d1 = torch.frombuffer(bytearray(g1), dtype=torch.uint8).to(dev)
d2 = torch.frombuffer(bytearray(g2), dtype=torch.uint8).to(dev)
out = torch.zeros(12 * curve.fp_bytes, dtype=torch.uint8, device=dev)
flag = torch.zeros(1, dtype=torch.int32, device=dev)
for _ in range(a.reps):
    ctx.pairing_product_dev(cid, d1.data_ptr(), d2.data_ptr(), a.pairs, out.data_ptr(), flag.data_ptr(), torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()
assert int(flag.item()) == 1
print("ok", a.pairs, "pairs")
