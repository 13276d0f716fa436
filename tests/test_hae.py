"""Hashed aggregation exponents (bgls/blsHAE.go): BLAKE2Xb on the host (CPU test pins its compression function and
parameter block against hashlib; parity with the Go XOF itself is unpinned -- no vector in the reference), and the
reference's own HAE tests (bgls/blsHAE_test.go:14-83) through the engine on a GPU, with the scaled keys / signatures
checked against the oracle."""
import random

import pytest


def test_blake2xb_structure():
    from bgls_b200 import blake2x
    assert blake2x._selfcheck()
    a = blake2x.blake2xb(b"keys", 16 * 5)
    assert len(a) == 80 and a != blake2x.blake2xb(b"keys", 16 * 6)[:80]    # the output length is part of the hash
    assert blake2x.blake2xb(b"keys", 200)[:64] != blake2x.blake2xb(b"keyz", 200)[:64]
    assert len(set(blake2x.blake2xb(b"", 64 * 3)[i:i + 64] for i in (0, 64, 128))) == 3   # node offsets differ


@pytest.mark.gpu
def test_hae_reference_tests():
    from bgls_b200 import bgls as B
    from bgls_b200.curves import AggregatePoints, Altbn128, Bls12
    from oracle import c_oracle as C
    rng = random.Random(0x4AE)
    for curve in (Altbn128, Bls12):
        # TestAggregationWithHAE (blsHAE_test.go:14-57)
        N = 5
        msgs, sigs, pubkeys = [], [], []
        for i in range(N):
            msgs.append(rng.randbytes(32))
            sk, vk, _ = B.KeyGen(curve, rng)
            sigs.append(B.Sign(curve, sk, msgs[i]))
            pubkeys.append(vk)
        agg = B.AggregateSignaturesWithHAE(sigs, pubkeys)
        t = B.hashPubKeysToExponents(pubkeys)
        assert all(0 <= x < 1 << 128 for x in t) and len(set(t)) == N
        # the engine's scaled-and-summed signature equals the oracle's
        scaled = C.scale_points(curve.cid, 1, b"".join(s.raw for s in sigs), b"".join(x.to_bytes(32, "big") for x in t), N, 1)
        assert agg.raw == C.aggregate(curve.cid, 1, scaled, N, 1)
        assert B.VerifyAggregateSignatureWithHAE(curve, agg, pubkeys, msgs)
        assert not B.VerifyAggregateSignatureWithHAE(curve, agg, pubkeys[:N - 1], msgs)
        assert B.AggregateSignaturesWithHAE(sigs, pubkeys[:N - 1]) is None
        skf, vkf, _ = B.KeyGen(curve, rng)
        pubkeys.append(vkf)
        msgs.append(msgs[0])
        sigs.append(B.Sign(curve, skf, msgs[N]))
        agg = B.AggregateSignaturesWithHAE(sigs, pubkeys)
        assert B.VerifyAggregateSignatureWithHAE(curve, agg, pubkeys, msgs), "HAE must accept duplicate messages"
        assert not B.VerifyAggregateSignatureWithHAE(curve, agg, pubkeys[:N], msgs[:N])
        msgs[0], msgs[1] = msgs[1], msgs[N]
        assert not B.VerifyAggregateSignatureWithHAE(curve, AggregatePoints(sigs[:N]), pubkeys[:N], msgs[:N])
        # TestMultiSigWithHAE (blsHAE_test.go:59-83)
        msg = rng.randbytes(32)
        signers, sigs = [], []
        for j in range(8):
            sk, vk, _ = B.KeyGen(curve, rng)
            sigs.append(B.Sign(curve, sk, msg))
            signers.append(vk)
        agg = B.AggregateSignaturesWithHAE(sigs, signers)
        assert B.VerifyMultiSignatureWithHAE(curve, agg, signers, msg)
        assert not B.VerifyMultiSignatureWithHAE(curve, agg, signers, rng.randbytes(32))
        signers[0] = B.KeyGen(curve, rng)[1]
        assert not B.VerifyMultiSignatureWithHAE(curve, agg, signers, msg)
