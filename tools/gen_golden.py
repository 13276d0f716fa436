"""Regenerates tests/golden/pairing_golden.json from the Python big-int oracle
(oracle/bgls_oracle.py).  Run from the repo root:  python tools/gen_golden.py
The reference holds no pairing known-answer vector, so these are the oracle's
own definition-level outputs (direct (p^12-1)/r power), used to pin the C oracle and the CUDA path."""
import json
import os
import random
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import bgls_oracle as O  # noqa: E402


def main():
    rng = random.Random(0xB615)
    out = {}
    for c in (O.ALTBN128, O.BLS12_381):
        n = 4
        ks = [rng.randrange(1, c.r) for _ in range(2 * n)]
        Ps = [c.g1_mul(c.g1, k) for k in ks[:n]]
        Qs = [c.g2_mul(c.g2, k) for k in ks[n:]]
        g1 = b"".join(c.marshal_g1(P) for P in Ps)
        g2 = b"".join(c.marshal_g2(Q) for Q in Qs)
        entry = {
            "g1": g1.hex(), "g2": g2.hex(), "n": n,
            "product_gt": c.marshal_gt(c.pairing_product(Ps, Qs)).hex(),
            "gen_gt": c.marshal_gt(c.pair(c.g1, c.g2)).hex(),
            "pair0_gt": c.marshal_gt(c.pair(Ps[0], Qs[0])).hex(),
            "sum_g1": c.marshal_g1(O.aggregate_points(c, Ps, "g1")).hex(),
            "sum_g2": c.marshal_g2(O.aggregate_points(c, Qs, "g2")).hex(),
        }
        # a valid 3-signer aggregate signature (bgls/bgls_test.go:40-57 shape)
        msgs = [bytes(rng.randrange(256) for _ in range(32)) for _ in range(3)]
        keys = [O.keygen(c, rng) for _ in range(3)]
        sig = O.aggregate_points(c, [O.sign(c, sk, m) for (sk, _), m in zip(keys, msgs)], "g1")
        assert O.verify_agg_sig(c, sig, [pk for _, pk in keys], msgs)
        entry["agg"] = {
            "msgs": [m.hex() for m in msgs],
            "hashes": [c.marshal_g1(c.hash_to_g1(m)).hex() for m in msgs],
            "pubkeys": [c.marshal_g2(pk) for _, pk in keys] and [c.marshal_g2(pk).hex() for _, pk in keys],
            "sig": c.marshal_g1(sig).hex(),
            "g2gen": c.marshal_g2(c.g2).hex(),
        }
        out[c.name] = entry
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "pairing_golden.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", path)


if __name__ == "__main__":
    main()
