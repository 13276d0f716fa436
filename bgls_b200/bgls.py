"""Scheme layer: a transliteration of the reference's bgls/bgls.go and the Kosk entry points of
bgls/blsKosk.go on top of the engine-backed curve mirror (bgls_b200/curves.py).  Function names,
argument order and accept/reject behaviour follow the reference line by line:

    KeyGen / LoadPublicKey          bgls/bgls.go:30-43
    Sign / SignCustHash             bgls/bgls.go:46-56
    VerifySingleSignature[CustHash] bgls/bgls.go:59-70
    VerifyAggregateSignature        bgls/bgls.go:82-84   -> verifyAggSig bgls/bgls.go:94-119
    verifyMultiSignature            bgls/bgls.go:89-92
    AggregateSignatures / Keys      bgls/bgls.go:123-131
    KoskSign / KoskVerify*          bgls/blsKosk.go:74-150 (incl. batch multi-signature and multiplicity)
    DistinctMsg*                    bgls/blsDistinctMessage.go:22-57
    *WithHAE                        bgls/blsHAE.go:39-93 (hashed aggregation exponents: BLAKE2Xb on the host, ScalePoints on the GPU)
    VerifyAggregateSignatures       engine extension: many independent aggregate verifies in one batched launch

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
    if hash is None and len(keys) >= 1 and all(curve._is(k, 2) for k in keys) and curve._is(aggsig, 1):
        # default hash: one engine call (bgls_verify_multi_signature)
        return curve._ctx().verify_multi_signature(curve.cid, bytes(msg), b"".join(k._canon() for k in keys), len(keys), aggsig._canon())
    vs = AggregatePoints(keys)
    return VerifySingleSignatureCustHash(curve, aggsig, vs, msg, hash or curve.HashToG1)


def verifyAggSig(curve: CurveSystem, aggsig: Point, keys, msgs, allowDuplicates: bool, hash=None) -> bool:
    if len(keys) != len(msgs):
        return False
    if not allowDuplicates and containsDuplicateMessage(msgs):
        return False
    if hash is None and all(curve._is(k, 2) for k in keys) and curve._is(aggsig, 1):
        # default hash: the whole function is one engine call (bgls_verify_aggregate_signature); the duplicate check
        # above is repeated there only when allowDuplicates is false, with the same verdict
        return curve._ctx().verify_aggregate_signature(curve.cid, [bytes(m) for m in msgs], b"".join(k._canon() for k in keys),
                                                       aggsig._canon(), allow_duplicates=True)
    hashed = curve.HashToG1Many(msgs) if hash is None else [hash(m) for m in msgs]
    pts1 = hashed + [aggsig.Mul(-1)]
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


def KoskVerifyBatchMultiSignature(curve, aggsigs, pubkeys, msgs):
    """bgls/blsKosk.go:126-133: one aggregate check over the per-message aggregated keys."""
    aggsig = AggregateSignatures(aggsigs)
    keys = [AggregateKeys(pk) for pk in pubkeys]
    return KoskVerifyAggregateSignature(curve, aggsig, keys, msgs)


def KoskVerifyMultiSignatureWithMultiplicity(curve, aggsig, keys, multiplicity, msg):
    """bgls/blsKosk.go:137-150: keys scaled by their multiplicities (ScalePoints), then the multi-signature check."""
    from .curves import ScalePoints
    if multiplicity is None:
        return KoskVerifyMultiSignature(curve, aggsig, keys, msg)
    if len(keys) != len(multiplicity):
        return False
    return KoskVerifyMultiSignature(curve, aggsig, ScalePoints(keys, [int(m) for m in multiplicity]), msg)


# ---- Distinct messages (bgls/blsDistinctMessage.go): the public key is prepended to the message
def DistinctMsgSign(curve, sk, m, hash=None):
    msg = LoadPublicKey(curve, sk).MarshalUncompressed() + bytes(m)
    return SignCustHash(sk, msg, hash or curve.HashToG1)


def DistinctMsgVerifySingleSignature(curve, sig, pubkey, m):
    return VerifySingleSignature(curve, sig, pubkey, pubkey.MarshalUncompressed() + bytes(m))


def DistinctMsgVerifyAggregateSignature(curve, aggsig, keys, msgs):
    if len(keys) != len(msgs):
        return False
    return verifyAggSig(curve, aggsig, keys, [k.MarshalUncompressed() + bytes(m) for k, m in zip(keys, msgs)], True)


# ---- Hashed aggregation exponents (bgls/blsHAE.go)
def hashPubKeysToExponents(pubkeys):
    """bgls/blsHAE.go:81-93: the uncompressed marshal of every key goes into one BLAKE2Xb instance of 16 n output
    bytes; t_i is the i-th 16-byte big-endian word."""
    from .blake2x import blake2xb
    n = len(pubkeys)
    if n == 0:
        return []
    out = blake2xb(b"".join(pk.MarshalUncompressed() for pk in pubkeys), 16 * n)
    return [int.from_bytes(out[16 * i:16 * i + 16], "big") for i in range(n)]


def AggregateSignaturesWithHAE(sigs, pubkeys):
    """bgls/blsHAE.go:39-46: nil when the counts differ."""
    from .curves import ScalePoints
    if len(pubkeys) != len(sigs):
        return None
    return AggregatePoints(ScalePoints(sigs, hashPubKeysToExponents(pubkeys)))


def VerifyAggregateSignatureWithHAE(curve, aggsig, pubkeys, msgs):
    """bgls/blsHAE.go:49-53: the keys scaled by their exponents (one ScalePoints launch), then verifyAggSig with
    duplicates allowed."""
    from .curves import ScalePoints
    return verifyAggSig(curve, aggsig, ScalePoints(pubkeys, hashPubKeysToExponents(pubkeys)), msgs, True)


def getAggregatePubKey(curve, pubkeys):
    """bgls/blsHAE.go:75-78"""
    from .curves import ScalePoints
    return AggregatePoints(ScalePoints(pubkeys, hashPubKeysToExponents(pubkeys)))


def VerifyMultiSignatureWithHAE(curve, aggsig, pubkeys, msg):
    """bgls/blsHAE.go:56-58"""
    return VerifySingleSignature(curve, aggsig, getAggregatePubKey(curve, pubkeys), msg)


def VerifyBatchMultiSignatureWithHAE(curve, aggsigs, aggpubkeys, msgs, allowDups):
    """bgls/blsHAE.go:62-72.  With allowDups the reference draws random factors and calls ScalePoints(aggsigs, t) but
    discards the result (blsHAE.go:68), so the factors never enter the check: the same happens here."""
    from .curves import ScalePoints
    if allowDups:
        ScalePoints(aggsigs, [secrets.randbelow(curve.GetG1Order()) for _ in aggsigs])
    return verifyAggSig(curve, AggregateSignatures(aggsigs), aggpubkeys, msgs, True)


def VerifyAggregateSignatures(curve, items):
    """Engine extension (BASELINE config 5, throughput mode): `items` = [(aggsig, keys, msgs), ...]; every entry gets
    the verdict VerifyAggregateSignature would give it, with all hashes in one launch and all pairing products in
    one batched launch."""
    verdicts, todo, flat = [None] * len(items), [], []
    for i, (aggsig, keys, msgs) in enumerate(items):
        if len(keys) != len(msgs) or containsDuplicateMessage(msgs):
            verdicts[i] = False
        else:
            todo.append(i)
            flat += list(msgs)
    hashed = curve.HashToG1Many(flat)
    products, pos = [], 0
    for i in todo:
        aggsig, keys, msgs = items[i]
        products.append((hashed[pos:pos + len(msgs)] + [aggsig.Mul(-1)], list(keys) + [curve.GetG2()]))
        pos += len(msgs)
    res = curve.PairingChecks(products)
    for j, i in enumerate(todo):
        verdicts[i] = bool(res[j]) if res is not None else False
    return verdicts
