"""Host-side mirror of the reference's curve abstraction (package `curves`), backed by the
CUDA engine through the C ABI.  Same names, argument meaning and error behaviour as

    curves/curve.go:12-70   CurveSystem / Point / PointT
    curves/curve.go:73-110  AggregatePoints        curves/curve.go:190-206  ScalePoints
    curves/altbn128.go      Altbn128 singleton     curves/bls12_381.go      Bls12 singleton

so that the scheme layer (bgls_b200/bgls.py, a transliteration of bgls/bgls.go) and the parity
tests read like the reference's own.  Go's `(value, ok)` returns become Python tuples.

Only glue lives here (byte packing, y -> p - y negation exactly as the reference's Negate does
through ToAffineCoords/MakePoint, curves/altbn128.go:123-128).  Every group / pairing operation
is a call into libbgls_b200.so; if the library or a GPU is missing these calls raise.
The Go toolchain is not available in the build image, so this mirror is Python; the cgo binding
a maintainer would add on the Go side is in INTEGRATION.md.
"""
from __future__ import annotations

import os
import threading

from . import _native
from ._native import ALTBN128, BLS12_381, BglsError

_ctx_lock = threading.Lock()
_ctxs = {}


def get_context(device: int | None = None) -> _native.Context:
    """Process-wide engine context for `device` (default: $BGLS_DEVICE, else $LOCAL_RANK, else 0)."""
    if device is None:
        device = int(os.environ.get("BGLS_DEVICE", os.environ.get("LOCAL_RANK", "0")))
    with _ctx_lock:
        if device not in _ctxs:
            _ctxs[device] = _native.Context(device)
        return _ctxs[device]


class Point:
    """curves/curve.go:52-60.  Holds the uncompressed affine record (MarshalUncompressed bytes)."""

    __slots__ = ("curve", "group", "raw")

    def __init__(self, curve: "CurveSystem", group: int, raw: bytes):
        self.curve, self.group, self.raw = curve, group, bytes(raw)

    def Add(self, other):
        """(Point, ok); ok is False on a type mismatch (curves/altbn128.go:59-66,181-188)."""
        if not isinstance(other, Point) or other.curve is not self.curve or other.group != self.group:
            return None, False
        out = self.curve._ctx().aggregate_points(self.curve.cid, self.group, self.raw + other.raw, 2)
        return Point(self.curve, self.group, out), True

    def Copy(self):
        return Point(self.curve, self.group, self.raw)

    def Equals(self, other) -> bool:
        return isinstance(other, Point) and other.curve is self.curve and other.group == self.group and self._canon() == other._canon()

    def _is_inf(self) -> bool:
        return not any(self.raw) or (self.curve.cid == BLS12_381 and bool(self.raw[0] & 0x40))

    def _canon(self) -> bytes:
        return bytes(len(self.raw)) if self._is_inf() else self.raw

    def MarshalUncompressed(self) -> bytes:
        return self._canon()

    def Marshal(self) -> bytes:
        """Compressed form, computed by the engine (curves/altbn128.go:81-89,203-221; bls12_381.go:57-59,118-120)."""
        return self.curve._ctx().compress_points(self.curve.cid, self.group, self._canon(), 1)

    def Negate(self):
        """curves/altbn128.go:123-128,227-233; bls12_381.go:85-91,139-146."""
        if self._is_inf():
            return self.Copy()
        F, q = self.curve.fp_bytes, self.curve.q
        half = len(self.raw) // 2
        ys = [int.from_bytes(self.raw[half + i:half + i + F], "big") for i in range(0, half, F)]
        neg = b"".join(((q - y) % q).to_bytes(F, "big") for y in ys)
        return Point(self.curve, self.group, self.raw[:half] + neg)

    def Mul(self, scalar: int):
        """curves/altbn128.go:107-121,235-249: negative scalars negate the point, zero gives infinity."""
        scalar = int(scalar)
        if scalar == 0:
            return self.curve.GetG1Infinity() if self.group == 1 else self.curve.GetG2Infinity()
        base = self
        if scalar < 0:
            base, scalar = self.Negate(), -scalar
        if scalar == 1:
            return base.Copy()
        k = (scalar % self.curve.order).to_bytes(32, "big")
        out = self.curve._ctx().scale_points(self.curve.cid, self.group, base._canon(), k, 1)
        return Point(self.curve, self.group, out)

    def ToAffineCoords(self):
        """[x, y] (G1) or [x_im, x_re, y_im, y_re] (G2): curves/altbn128.go:147-155,251-262."""
        F = self.curve.fp_bytes
        r = self._canon()
        return [int.from_bytes(r[i:i + F], "big") for i in range(0, len(r), F)]


class PointT:
    """curves/curve.go:63-70; GT element as its 12*F-byte marshal form."""

    __slots__ = ("curve", "raw")

    def __init__(self, curve, raw: bytes):
        self.curve, self.raw = curve, bytes(raw)

    def Add(self, other):
        """GT multiplication (curves/altbn128.go:264-271; bls12_381.go:160-168)."""
        if not isinstance(other, PointT) or other.curve is not self.curve:
            return None, False
        return PointT(self.curve, self.curve._ctx().gt_mul(self.curve.cid, self.raw, other.raw)), True

    def Copy(self):
        return PointT(self.curve, self.raw)

    def Equals(self, other) -> bool:
        return isinstance(other, PointT) and other.curve is self.curve and self.raw == other.raw

    def Marshal(self) -> bytes:
        return self.raw

    def Mul(self, scalar):
        """GT exponentiation (curves/altbn128.go:290-294; bls12_381.go:186-195) on the engine (bgls_gt_pow)."""
        e = int(scalar)
        if abs(e) >> 256:
            e %= self.curve.order      # GT has order r
        return PointT(self.curve, self.curve._ctx().gt_pow(self.curve.cid, self.raw, e))


class CurveSystem:
    """curves/curve.go:12-49."""

    def __init__(self, name, cid, q, order, g1, g2):
        self.name, self.cid, self.q, self.order = name, cid, q, order
        self.fp_bytes = _native.FP_BYTES[cid]
        self._g1 = b"".join(v.to_bytes(self.fp_bytes, "big") for v in g1)
        self._g2 = b"".join(v.to_bytes(self.fp_bytes, "big") for v in g2)
        self._gt = None
        self.device = None

    def _ctx(self):
        return get_context(self.device)

    def Name(self):
        return self.name

    # ---- constructors
    def _validated(self, group, raw, always):
        """The checks the reference makes before a Point exists (bgls_validate_points, mode REFERENCE): altbn128 builds
        every point through bn256.Unmarshal (on the curve; G2 also in the order-r subgroup; `check` is ignored,
        curves/altbn128.go:39-57,160-179), bls12-381 calls Check() when asked to (curves/bls12_381.go:197-226) and on every
        Unmarshal (:242-264)."""
        if always and not self._ctx().validate_points(self.cid, group, raw, 1, reference=True)[0]:
            return None, False
        return Point(self, group, raw), True

    def MakeG1Point(self, coords, check=True):
        if len(coords) != 2 or any(not (0 <= int(c) < self.q) for c in coords):
            return None, False
        raw = b"".join(int(c).to_bytes(self.fp_bytes, "big") for c in coords)
        return self._validated(1, raw, self.cid == ALTBN128 or check)

    def MakeG2Point(self, coords, check=True):
        if len(coords) != 4 or any(not (0 <= int(c) < self.q) for c in coords):
            return None, False
        raw = b"".join(int(c).to_bytes(self.fp_bytes, "big") for c in coords)
        return self._validated(2, raw, self.cid == ALTBN128 or check)

    def _unmarshal(self, group, data):
        """curves/altbn128.go:296-376, bls12_381.go:242-264: uncompressed or compressed by length; the compressed form
        is decoded on the GPU (bls12-381 with the subgroup check the reference's Check() performs)."""
        F = self.fp_bytes
        if data is None:
            return None, False
        data = bytes(data)
        if len(data) == 2 * group * F:
            return self._validated(group, data, True)
        if len(data) != group * F:
            return None, False
        # subgroup membership: bls12-381 Check(), and the upstream twist check of altbn128 G2
        raw, ok = self._ctx().decompress_points(self.cid, group, data, 1, check_subgroup=(self.cid == BLS12_381 or group == 2))
        return (Point(self, group, raw), True) if ok[0] else (None, False)

    def UnmarshalG1(self, data):
        return self._unmarshal(1, data)

    def UnmarshalG2(self, data):
        return self._unmarshal(2, data)

    def UnmarshalGT(self, data):
        if data is None or len(data) != 12 * self.fp_bytes:
            return None, False
        return PointT(self, data), True

    def GetG1(self):
        return Point(self, 1, self._g1)

    def GetG2(self):
        return Point(self, 2, self._g2)

    def GetG1Infinity(self):
        return Point(self, 1, bytes(2 * self.fp_bytes))

    def GetG2Infinity(self):
        return Point(self, 2, bytes(4 * self.fp_bytes))

    def GetGTIdentity(self):
        """Pair(G1, inf) in the reference (curves/altbn128.go:478): the element 1."""
        F = self.fp_bytes
        return PointT(self, bytes(12 * F - 1) + b"\x01")

    def GetGT(self):
        if self._gt is None:
            self._gt, _ = self.Pair(self.GetG1(), self.GetG2())
        return self._gt

    def GetG1Q(self):
        return self.q

    def GetG1Order(self):
        return self.order

    def HashToG1(self, message: bytes):
        """curves/altbn128.go:509-513 (Keccak-256 try-and-increment), curves/bls12_381.go:349-351
        (blake2b + Fouque-Tibouchi): computed by the engine's hash-to-G1 kernel."""
        out = self._ctx().hash_to_g1(self.cid, [bytes(message)])
        return Point(self, 1, out)

    def HashToG1Many(self, messages):
        """The n concurrentHash goroutines of verifyAggSig (bgls/bgls.go:106-111,134-137) as ONE kernel launch."""
        msgs = [bytes(m) for m in messages]
        if not msgs:
            return []
        out = self._ctx().hash_to_g1(self.cid, msgs)
        rec = 2 * self.fp_bytes
        return [Point(self, 1, out[i * rec:(i + 1) * rec]) for i in range(len(msgs))]

    def PairingChecks(self, products):
        """Engine extension for throughput mode (BASELINE config 5): `products` is a list of (pts1, pts2) pairs of
        equal-length Point lists; returns [prod_i e(pts1[i], pts2[i]) == 1] for each, all in one batched launch
        (bgls_pairing_check_batch).  None when a product is malformed (the reference's (nil, false))."""
        g1, g2, offs = [], [], [0]
        for pts1, pts2 in products:
            if len(pts1) != len(pts2) or any(not self._is(a, 1) for a in pts1) or any(not self._is(b, 2) for b in pts2):
                return None
            g1 += [p._canon() for p in pts1]
            g2 += [p._canon() for p in pts2]
            offs.append(offs[-1] + len(pts1))
        if not products:
            return []
        return self._ctx().pairing_check_batch(self.cid, b"".join(g1), b"".join(g2), offs)

    # ---- the accelerated boundary
    def Pair(self, p1, p2):
        """curves/altbn128.go:130-141; bls12_381.go:228-236: (PointT, ok), ok False on a type mismatch."""
        if not self._is(p1, 1) or not self._is(p2, 2):
            return None, False
        return PointT(self, self._ctx().pair(self.cid, p1._canon(), p2._canon())), True

    def PairingProduct(self, pts1, pts2):
        """curves/curve.go:125-170 via altbn128.go:143-145 / bls12_381.go:238-240."""
        if len(pts1) != len(pts2):
            return None, False
        for a, b in zip(pts1, pts2):
            if not self._is(a, 1) or not self._is(b, 2):
                return None, False
        g1 = b"".join(p._canon() for p in pts1)
        g2 = b"".join(p._canon() for p in pts2)
        gt, _ = self._ctx().pairing_product(self.cid, g1, g2, len(pts1))
        return PointT(self, gt), True

    def _is(self, p, group):
        return isinstance(p, Point) and p.curve is self and p.group == group


def AggregatePoints(points):
    """curves/curve.go:73-110.  One engine call for the whole list (the reference's goroutine tree
    computes the same group sum).  len 1 returns the point itself; len 0 hangs in the reference
    (curve.go:94-108) and raises here."""
    if len(points) == 0:
        raise ValueError("AggregatePoints of an empty list does not terminate in the reference (curves/curve.go:94-108)")
    if len(points) == 1:
        return points[0]
    first = points[0]
    if any(not isinstance(p, Point) or p.curve is not first.curve or p.group != first.group for p in points):
        return None
    raw = b"".join(p._canon() for p in points)
    out = first.curve._ctx().aggregate_points(first.curve.cid, first.group, raw, len(points))
    return Point(first.curve, first.group, out)


def ScalePoints(pts, factors):
    """curves/curve.go:190-214: nil factors returns pts, a length mismatch returns nil, a nil factor copies."""
    if factors is None:
        return pts
    if len(pts) != len(factors):
        return None
    if not pts:
        return []
    first = pts[0]
    base, ks = [], []
    for p, f in zip(pts, factors):
        f = 1 if f is None else int(f)
        if f < 0:
            p, f = p.Negate(), -f
        base.append(p._canon())
        ks.append((f % first.curve.order).to_bytes(32, "big") if f else bytes(32))
    out = first.curve._ctx().scale_points(first.curve.cid, first.group, b"".join(base), b"".join(ks), len(pts))
    rec = len(first.raw)
    return [Point(first.curve, first.group, out[i * rec:(i + 1) * rec]) for i in range(len(pts))]


# curve constants: curves/altbn128.go:458-480, curves/altbn128_test.go:26-35, curves/bls12_381.go:328-346
Altbn128 = CurveSystem(
    "altbn128", ALTBN128,
    21888242871839275222246405745257275088696311157297823662689037894645226208583,
    21888242871839275222246405745257275088548364400416034343698204186575808495617,
    (1, 2),
    (11559732032986387107991004021392285783925812861821192530917403151452391805634,
     10857046999023057135944570762232829481370756359578518086990519993285655852781,
     4082367875863433681332203403145435568316851327593401208105741076214120093531,
     8495653923123431417604973247489272438418190587263600148770280649306958101930),
)
Bls12 = CurveSystem(
    "bls12", BLS12_381,
    0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab,
    52435875175126190479447740508185965837690552500527637822603658699938581184513,
    (0x17f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac586c55e83ff97a1aeffb3af00adb22c6bb,
     0x08b3f481e3aaa0f1a09e30ed741d8ae4fcf5e095d5d00af600db18cb2c04b3edd03cc744a2888ae40caa232946c5e7e1),
    (0x13e02b6052719f607dacd3a088274f65596bd0d09920b61ab5da61bbdc7f5049334cf11213945d57e5ac7d055d042b7e,
     0x024aa2b2f08f0a91260805272dc51051c6e47ad4fa403b02b4510b647ae3d1770bac0326a805bbefd48056c8c121bdb8,
     0x0606c4a02ea734cc32acd2b02bc28b99cb3e287e85a763af267492ab572e99ab3f370d275cec1da1aaa9075ff05f79be,
     0x0ce5d527727d6e118cc9cdc6da2e351aadfd9baa8cbdd3a76d429a695160d12c923ac9cc3baca289e193548608b82801),
)

__all__ = ["CurveSystem", "Point", "PointT", "AggregatePoints", "ScalePoints", "Altbn128", "Bls12", "get_context", "BglsError"]
