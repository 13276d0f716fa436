"""Scheme layer: a transliteration of the reference's bgls/bgls.go and the Kosk entry points of
bgls/blsKosk.go on top of the engine-backed curve mirror (bgls_b200/curves.py).  Function names,
argument order and accept/reject behaviour follow the reference line by line:

    KeyGen / LoadPublicKey          bgls/bgls.go:30-43
    Sign / SignCustHash             bgls/bgls.go:46-56
    VerifySingleSignature[CustHash] bgls/bgls.go:59-70
    VerifyAggregateSignature        bgls/bgls.go:82-84   -> verifyAggSig bgls/bgls.go:94-119
    verifyMultiSignature            bgls/bgls.go:89-92
    AggregateSignatures / Keys      bgls/bgls.go:123-131
    KoskSign / KoskVerify*          bgls/blsKosk.go:74-120

The only computation done here is control flow; every pairing, point sum, scalar multiplication and
hash-to-G1 goes through the CurveSystem (i.e. the CUDA engine).  `hash` parameters mirror the
reference's *CustHash variants (a callable msg -> Point).
"""
from __future__ import annotations

import secrets

from .curves import AggregatePoints, CurveSystem, Point


def KeyGen(curve: CurveSystem, rng=None):
    x = (rng.randrange(curve.GetG1Order()) if rng is not None else secrets.randbelow(curve.GetG1Order()))
    return x, LoadPublicKey(curve, x), None


def LoadPublicKey(curve: CurveSystem, sk: int) -> Point:
    return curve.GetG2().Mul(sk)


def Sign(curve: CurveSystem, sk: int, msg: bytes) -> Point:
    return SignCustHash(sk, msg, curve.HashToG1)


def SignCustHash(sk: int, msg: bytes, hash) -> Point:
    return hash(msg).Mul(sk)


def VerifySingleSignature(curve: CurveSystem, sig: Point, pubKey: Point, msg: bytes) -> bool:
    return VerifySingleSignatureCustHash(curve, sig, pubKey, msg, curve.HashToG1)


def VerifySingleSignatureCustHash(curve: CurveSystem, sig: Point, pubkey: Point, msg: bytes, hash) -> bool:
    h = hash(msg).Mul(-1)
    paired, _ = curve.PairingProduct([h, sig], [pubkey, curve.GetG2()])
    return curve.GetGTIdentity().Equals(paired)


def VerifyAggregateSignature(curve: CurveSystem, aggsig: Point, keys, msgs) -> bool:
    return verifyAggSig(curve, aggsig, keys, msgs, False)


def verifyMultiSignature(curve: CurveSystem, aggsig: Point, keys, msg: bytes, hash=None) -> bool:
    vs = AggregatePoints(keys)
    return VerifySingleSignatureCustHash(curve, aggsig, vs, msg, hash or curve.HashToG1)


def verifyAggSig(curve: CurveSystem, aggsig: Point, keys, msgs, allowDuplicates: bool, hash=None) -> bool:
    if len(keys) != len(msgs):
        return False
    if not allowDuplicates and containsDuplicateMessage(msgs):
        return False
    hash = hash or curve.HashToG1
    pts1 = [hash(m) for m in msgs] + [aggsig.Mul(-1)]
    pts2 = list(keys) + [curve.GetG2()]
    aggPt, ok = curve.PairingProduct(pts1, pts2)
    if ok:
        return aggPt.Equals(curve.GetGTIdentity())
    return ok


def AggregateSignatures(sigs):
    return AggregatePoints(sigs)


def AggregateKeys(keys):
    return AggregatePoints(keys)


def containsDuplicateMessage(msgs) -> bool:
    seen = set()
    for m in msgs:
        m = bytes(m)
        if m in seen:
            return True
        seen.add(m)
    return False


# ---- Kosk (bgls/blsKosk.go): 0x01 domain-separation prefix
def KoskSign(curve, sk, msg, hash=None):
    return SignCustHash(sk, b"\x01" + bytes(msg), hash or curve.HashToG1)


def KoskVerifySingleSignature(curve, sig, pubKey, msg, hash=None):
    return VerifySingleSignatureCustHash(curve, sig, pubKey, b"\x01" + bytes(msg), hash or curve.HashToG1)


def KoskVerifyAggregateSignature(curve, aggsig, keys, msgs, hash=None):
    return verifyAggSig(curve, aggsig, keys, [b"\x01" + bytes(m) for m in msgs], True, hash)


def KoskVerifyMultiSignature(curve, aggsig, keys, msg, hash=None):
    return verifyMultiSignature(curve, aggsig, keys, b"\x01" + bytes(msg), hash)
