"""Writes tests/golden/compressed_golden.json: compressed G1/G2 records of seeded points with the oracle's
restatement of the reference codec (curves/altbn128.go:81-89,203-221,296-376; bls12-381: zcash format of the
upstream library, unpinned).  python tools/gen_golden_compressed.py"""
import json
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import bgls_oracle as O  # noqa: E402

out = {}
for c in (O.ALTBN128, O.BLS12_381):
    rng = random.Random(2024)
    g1 = [c.g1_mul(c.g1, rng.randrange(1, c.r)) for _ in range(6)] + [None, c.g1, c.g1_neg(c.g1)]
    g2 = [c.g2_mul(c.g2, rng.randrange(1, c.r)) for _ in range(6)] + [None, c.g2, c.g2_neg(c.g2)]
    out[c.name] = {
        "g1": [{"uncompressed": c.marshal_g1(P).hex(), "compressed": c.compress_g1(P).hex()} for P in g1],
        "g2": [{"uncompressed": c.marshal_g2(Q).hex(), "compressed": c.compress_g2(Q).hex()} for Q in g2],
    }
with open(os.path.join(ROOT, "tests", "golden", "compressed_golden.json"), "w") as f:
    json.dump(out, f, indent=1)
print("wrote compressed_golden.json")
