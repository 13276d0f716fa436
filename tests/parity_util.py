"""Shared helpers for the parity tests: seeded synthetic inputs built with the C oracle."""
import json
import os
import random

from oracle import bgls_oracle as O
from oracle import c_oracle as C

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "pairing_golden.json")))
CURVES = [(0, O.ALTBN128), (1, O.BLS12_381)]


def rand_scalars(rng, c, n):
    return [rng.randrange(1, c.r) for _ in range(n)]


def scalars_bytes(ks):
    return b"".join(int(k).to_bytes(32, "big") for k in ks)


def rand_points(cid, c, group, n, rng, nthreads=8):
    """n pseudo-random points k_i * G (C oracle scalar multiplication)."""
    gen = c.marshal_g1(c.g1) if group == 1 else c.marshal_g2(c.g2)
    ks = rand_scalars(rng, c, n)
    return C.scale_points(cid, group, gen * n, scalars_bytes(ks), n, nthreads), ks


def make_aggregate(cid, c, n, rng, nthreads=8):
    """Synthetic n-signer aggregate signature with known discrete logs:
    H_i = h_i*G1, pk_i = s_i*G2, sigma = (sum h_i s_i)*G1.  Returns the packed (n+1)-pair arrays
    exactly as verifyAggSig builds them (bgls/bgls.go:103-113): pts1 = [H_0..H_{n-1}, -sigma],
    pts2 = [pk_0..pk_{n-1}, g2]."""
    hs, ss = rand_scalars(rng, c, n), rand_scalars(rng, c, n)
    H = C.scale_points(cid, 1, c.marshal_g1(c.g1) * n, scalars_bytes(hs), n, nthreads)
    PK = C.scale_points(cid, 2, c.marshal_g2(c.g2) * n, scalars_bytes(ss), n, nthreads)
    tot = sum(h * s for h, s in zip(hs, ss)) % c.r
    neg_sigma = C.scale_points(cid, 1, c.marshal_g1(c.g1), scalars_bytes([(c.r - tot) % c.r]), 1)
    return H + neg_sigma, PK + c.marshal_g2(c.g2)
