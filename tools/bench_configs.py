#!/usr/bin/env python
"""Times the BASELINE.json configs that are not the bench.py headline (profiles/ evidence, one GPU):
  config 3  bls12-381 MultiSignature verify, 65,536 signers: AggregatePoints over the G2 keys + the 2-pairing check
  config 5  batch of independent altbn128 aggregate verifies x 256 signers (per-GPU share of 4096: --checks)
Inputs are device resident, CUDA events on the launching stream, best of --reps after a warm-up.
    python tools/bench_configs.py [--checks 512] [--reps 5]"""
import argparse
import json
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--checks", type=int, default=512)
    ap.add_argument("--signers", type=int, default=256)
    ap.add_argument("--keys", type=int, default=65536)
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    import torch

    import bgls_b200
    from bgls_b200.curves import Altbn128, Bls12
    dev = torch.device("cuda", 0)
    ctx = bgls_b200.Context(0)
    rng = random.Random(5)
    to_b = lambda ks: b"".join(int(k).to_bytes(32, "big") for k in ks)

    def timed(fn):
        fn()
        torch.cuda.synchronize()
        best = 1e30
        for _ in range(args.reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        return best
    s = torch.cuda.current_stream().cuda_stream
    out = {}

    # ---- config 3
    cid, curve, r = 1, Bls12, 52435875175126190479447740508185965837690552500527637822603658699938581184513
    n = args.keys
    ks = [rng.randrange(1, r) for _ in range(n)]
    keys = ctx.scale_points(cid, 2, curve.GetG2().raw * n, to_b(ks), n)
    d_keys = torch.frombuffer(bytearray(keys), dtype=torch.uint8).to(dev)
    d_sum = torch.zeros(192, dtype=torch.uint8, device=dev)
    ms = timed(lambda: ctx.aggregate_points_dev(cid, 2, d_keys.data_ptr(), n, d_sum.data_ptr(), s))
    want = ctx.scale_points(cid, 2, curve.GetG2().raw, to_b([sum(ks) % r]), 1)
    assert bytes(d_sum.cpu().numpy()) == want, "aggregate differs from (sum k_i) G2"
    out["config3_g2_aggregate"] = {"keys": n, "ms": ms, "keys_per_s": n / (ms * 1e-3), "GB_per_s": n * 192 / (ms * 1e-3) / 1e9}
    # the whole verifyMultiSignature (bgls/bgls.go:89-92) through bgls_verify_multi_signature: pinned host buffers in, verdict out
    import ctypes
    import time
    msg = b"\x01" + bytes(range(64))
    H = ctx.hash_to_g1(cid, [msg])
    sig = ctx.scale_points(cid, 1, H, to_b([sum(ks) % r]), 1)
    h_keys = torch.frombuffer(bytearray(keys), dtype=torch.uint8).pin_memory()
    ok = ctypes.c_int(0)

    def multisig():
        rc = ctx._L.bgls_verify_multi_signature(ctx._h, cid, msg, len(msg), ctypes.c_char_p(h_keys.data_ptr()), n, sig, ctypes.byref(ok))
        assert rc == 0 and ok.value == 1
    multisig()
    t0 = time.perf_counter()
    for _ in range(args.reps):
        multisig()
    dt = (time.perf_counter() - t0) / args.reps
    out["config3_verify_multi_signature_e2e"] = {"keys": n, "ms": dt * 1e3, "h2d_bytes": len(keys) + len(msg) + len(sig),
                                                 "api": "bgls_verify_multi_signature (pinned host buffers)"}

    # ---- config 5
    cid, curve, r = 0, Altbn128, 21888242871839275222246405745257275088548364400416034343698204186575808495617
    S, B = args.signers, args.checks
    hs, ss = [rng.randrange(1, r) for _ in range(S)], [rng.randrange(1, r) for _ in range(S)]
    g1 = ctx.scale_points(cid, 1, curve.GetG1().raw * S, to_b(hs), S)
    g2 = ctx.scale_points(cid, 2, curve.GetG2().raw * S, to_b(ss), S)
    tot = sum(h * x for h, x in zip(hs, ss)) % r
    neg = ctx.scale_points(cid, 1, curve.GetG1().raw, to_b([(r - tot) % r]), 1)
    a1, a2 = g1 + neg, g2 + curve.GetG2().raw
    bad1 = g1 + curve.GetG1().raw
    expect = [(b % 7) != 3 for b in range(B)]
    G1 = torch.frombuffer(bytearray(b"".join(a1 if e else bad1 for e in expect)), dtype=torch.uint8).to(dev)
    G2 = torch.frombuffer(bytearray(a2 * B), dtype=torch.uint8).to(dev)
    off = torch.arange(0, (B + 1) * (S + 1), S + 1, dtype=torch.int64, device=dev)
    ok = torch.zeros(B, dtype=torch.uint8, device=dev)
    total = B * (S + 1)
    ms = timed(lambda: ctx.pairing_check_batch_dev(cid, G1.data_ptr(), G2.data_ptr(), off.data_ptr(), B, total, ok.data_ptr(), s))
    assert [bool(x) for x in ok.cpu().tolist()] == expect, "batch verdicts differ"
    out["config5_batch"] = {"checks": B, "pairs_per_check": S + 1, "ms": ms, "pairings_per_s": total / (ms * 1e-3), "checks_per_s": B / (ms * 1e-3)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
