"""TEST INFRASTRUCTURE (oracle) -- not part of the product path.

Big-int CPU restatement of the Project-Arda/bgls hot path, written for obviousness, not
speed (about 0.5-1 s per pairing).  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline leg may import this.

What is restated, with the reference lines it follows (/root/reference/...):
  * curve constants                       curves/altbn128.go:458-480, curves/bls12_381.go:328-346,
                                          curves/README.md:9-19, curves/altbn128_test.go:26-35
  * hash-to-G1 (Keccak try&increment)     curves/hash.go:53-77, curves/altbn128.go:494-522
  * hash-to-G1 (blake2b + SvdW + cofactor) curves/hash.go:86-190, curves/bls12_381.go:349-400
  * byte layouts                          curves/altbn128.go:42-57,149-179,253-262; bls12_381.go:147-158
  * Pair / PairingProduct                 curves/curve.go:125-170,217-223; altbn128.go:130-145; bls12_381.go:228-240
  * AggregatePoints                       curves/curve.go:73-121
  * scheme layer                          bgls/bgls.go:30-150, bgls/blsKosk.go:74-120

The pairing arithmetic itself lives in third-party Go modules that are NOT vendored in the
reference (github.com/ethereum/go-ethereum/crypto/bn256/cloudflare and github.com/dis2/bls12,
both un-pinned).  It is restated here from the published definition:
    e(P, Q) = f_{lambda,Q}(P) ^ ((p^12 - 1) / r)          (optimal ate, reduced)
with Fp2 = Fp[i]/(i^2+1), Fp12 = Fp2[w]/(w^6 - xi):
    altbn128 : xi = 9+i, D-type twist y^2 = x^3 + 3/xi,  lambda = 6u+2 (+ two Frobenius lines)
    bls12-381: xi = 1+i, M-type twist y^2 = x^3 + 4*xi,  lambda = |x|, conjugate because x < 0
PARITY STATUS: hash-to-G1, G1/G2 points, marshal layouts and verify booleans are pinned by the
reference's own known-answer vectors (tests/test_oracle_kats.py).  GT *bytes* are
"parity unpinned": the reference holds no pairing known-answer vector; they are pinned only by
the definition above plus bilinearity / non-degeneracy / order-r properties.
"""
from __future__ import annotations

import hashlib

from .keccak import keccak256

# --------------------------------------------------------------------------------------
# Fp2 helpers: element = (re, im) meaning re + im*i, i^2 = -1
# --------------------------------------------------------------------------------------


class Fp:
    def __init__(self, p):
        self.p = p
        self.zero, self.one = 0, 1

    def add(self, a, b):
        return (a + b) % self.p

    def sub(self, a, b):
        return (a - b) % self.p

    def neg(self, a):
        return (-a) % self.p

    def mul(self, a, b):
        return (a * b) % self.p

    def inv(self, a):
        return pow(a, -1, self.p)

    def small(self, k):
        return k % self.p


class Fp2:
    def __init__(self, p):
        self.p = p
        self.zero, self.one = (0, 0), (1, 0)

    def add(self, a, b):
        return ((a[0] + b[0]) % self.p, (a[1] + b[1]) % self.p)

    def sub(self, a, b):
        return ((a[0] - b[0]) % self.p, (a[1] - b[1]) % self.p)

    def neg(self, a):
        return ((-a[0]) % self.p, (-a[1]) % self.p)

    def mul(self, a, b):
        return ((a[0] * b[0] - a[1] * b[1]) % self.p, (a[0] * b[1] + a[1] * b[0]) % self.p)

    def inv(self, a):
        n = pow(a[0] * a[0] + a[1] * a[1], -1, self.p)
        return (a[0] * n % self.p, (-a[1]) * n % self.p)

    def conj(self, a):
        return (a[0], (-a[1]) % self.p)

    def small(self, k):
        return (k % self.p, 0)

    def pow(self, a, e):
        r = self.one
        while e:
            if e & 1:
                r = self.mul(r, a)
            a = self.mul(a, a)
            e >>= 1
        return r


# --------------------------------------------------------------------------------------
# generic short-Weierstrass (a = 0) affine group law; None is the point at infinity
# --------------------------------------------------------------------------------------


def ec_add(F, P, Q):
    if P is None:
        return Q
    if Q is None:
        return P
    if P[0] == Q[0]:
        if P[1] != Q[1] or P[1] == F.zero:
            return None
        lam = F.mul(F.mul(F.small(3), F.mul(P[0], P[0])), F.inv(F.add(P[1], P[1])))
    else:
        lam = F.mul(F.sub(Q[1], P[1]), F.inv(F.sub(Q[0], P[0])))
    x3 = F.sub(F.sub(F.mul(lam, lam), P[0]), Q[0])
    y3 = F.sub(F.mul(lam, F.sub(P[0], x3)), P[1])
    return (x3, y3)


def ec_neg(F, P):
    return None if P is None else (P[0], F.neg(P[1]))


def ec_mul(F, P, k):
    if k < 0:
        return ec_mul(F, ec_neg(F, P), -k)
    R = None
    while k:
        if k & 1:
            R = ec_add(F, R, P)
        P = ec_add(F, P, P)
        k >>= 1
    return R


# --------------------------------------------------------------------------------------
# Curve
# --------------------------------------------------------------------------------------


class Curve:
    def __init__(self, name, p, r, b, xi, twist, g1, g2, nbytes, cofactor):
        self.name, self.p, self.r, self.b, self.xi, self.twist = name, p, r, b, xi, twist
        self.F1, self.F2 = Fp(p), Fp2(p)
        self.nbytes = nbytes  # bytes per Fp element at the boundary
        self.g1, self.g2 = g1, g2
        self.cofactor = cofactor
        if twist == "D":
            self.b2 = self.F2.mul((b, 0), self.F2.inv(xi))
        else:
            self.b2 = self.F2.mul((b, 0), xi)
        self.fp12_one = [(1, 0)] + [(0, 0)] * 5
        # w^-1 = w^5 / xi  (w^6 = xi)
        xi_inv = self.F2.inv(xi)
        self.w_inv = [(0, 0)] * 5 + [xi_inv]
        self.w_inv3 = self.fp12_mul(self.fp12_mul(self.w_inv, self.w_inv), self.w_inv)

    # ---- on-curve checks -------------------------------------------------------------
    def g1_on_curve(self, P):
        if P is None:
            return True
        x, y = P
        return (y * y - x * x * x - self.b) % self.p == 0

    def g2_on_curve(self, Q):
        if Q is None:
            return True
        F = self.F2
        x, y = Q
        return F.sub(F.mul(y, y), F.add(F.mul(x, F.mul(x, x)), self.b2)) == (0, 0)

    # ---- group law wrappers ----------------------------------------------------------
    def g1_add(self, P, Q):
        return ec_add(self.F1, P, Q)

    def g1_neg(self, P):
        return ec_neg(self.F1, P)

    def g1_mul(self, P, k):
        return ec_mul(self.F1, P, k)

    def g2_add(self, P, Q):
        return ec_add(self.F2, P, Q)

    def g2_neg(self, P):
        return ec_neg(self.F2, P)

    def g2_mul(self, P, k):
        return ec_mul(self.F2, P, k)

    # ---- Fp12 = Fp2[w]/(w^6 - xi): list of 6 Fp2 coefficients ---------------------------
    def fp12_mul(self, a, b):
        F = self.F2
        t = [(0, 0)] * 11
        for i in range(6):
            if a[i] == (0, 0):
                continue
            for j in range(6):
                if b[j] == (0, 0):
                    continue
                t[i + j] = F.add(t[i + j], F.mul(a[i], b[j]))
        out = list(t[:6])
        for k in range(6, 11):
            out[k - 6] = F.add(out[k - 6], F.mul(t[k], self.xi))
        return out

    def fp12_pow(self, a, e):
        r = self.fp12_one
        for bit in bin(e)[2:]:
            r = self.fp12_mul(r, r)
            if bit == "1":
                r = self.fp12_mul(r, a)
        return r

    def fp12_conj(self, a):
        """a^(p^6): w -> -w."""
        F = self.F2
        return [a[k] if k % 2 == 0 else F.neg(a[k]) for k in range(6)]

    # ---- pairing ---------------------------------------------------------------------
    def _line(self, T, R, P):
        """Line through twist points T,R (affine Fp2), untwisted to E(Fp12), evaluated at P in G1.
        Returns (line as Fp12, T+R).  A vertical line lies in a proper subfield and is
        annihilated by the final exponentiation, so it is returned as 1."""
        F = self.F2
        if T is None or R is None:
            return self.fp12_one, (R if T is None else T)
        xT, yT = T
        if T[0] == R[0]:
            if T[1] != R[1] or T[1] == (0, 0):
                return self.fp12_one, None
            lam = F.mul(F.mul((3, 0), F.mul(xT, xT)), F.inv(F.add(yT, yT)))
        else:
            lam = F.mul(F.sub(R[1], yT), F.inv(F.sub(R[0], xT)))
        x3 = F.sub(F.sub(F.mul(lam, lam), xT), R[0])
        y3 = F.sub(F.mul(lam, F.sub(xT, x3)), yT)
        xP, yP = P
        c_y = (yP, 0)
        c_x = F.neg(F.mul(lam, (xP, 0)))
        c_0 = F.sub(F.mul(lam, xT), yT)
        z = (0, 0)
        if self.twist == "D":
            # untwist (x w^2, y w^3), slope lam*w:  yP - lam xP w + (lam xT - yT) w^3
            line = [c_y, c_x, z, c_0, z, z]
        else:
            # untwist (x / w^2, y / w^3), slope lam/w: yP - lam xP w^-1 + (lam xT - yT) w^-3
            l1 = self.fp12_mul([c_x, z, z, z, z, z], self.w_inv)
            l3 = self.fp12_mul([c_0, z, z, z, z, z], self.w_inv3)
            line = [F.add(F.add(a, b), c) for a, b, c in zip([c_y, z, z, z, z, z], l1, l3)]
        return line, (x3, y3)

    def miller(self, P, Q):
        """f_{lambda,Q}(P) before final exponentiation; 1 if either input is infinity
        (reference: GT identity is *defined* as Pair(G1, inf) / Pair(inf, G2),
        curves/altbn128.go:478, curves/bls12_381.go:341)."""
        if P is None or Q is None:
            return self.fp12_one
        F = self.F2
        f = self.fp12_one
        T = Q
        if self.name == "altbn128":
            u = 4965661367192848881
            s = 6 * u + 2
            for bit in bin(s)[3:]:
                line, T = self._line(T, T, P)
                f = self.fp12_mul(self.fp12_mul(f, f), line)
                if bit == "1":
                    line, T = self._line(T, Q, P)
                    f = self.fp12_mul(f, line)
            p = self.p
            g2 = F.pow(self.xi, (p - 1) // 3)
            g3 = F.pow(self.xi, (p - 1) // 2)
            Q1 = (F.mul(F.conj(Q[0]), g2), F.mul(F.conj(Q[1]), g3))
            h2 = F.pow(self.xi, (p * p - 1) // 3)
            h3 = F.pow(self.xi, (p * p - 1) // 2)
            Q2 = (F.mul(Q[0], h2), F.mul(Q[1], h3))
            line, T = self._line(T, Q1, P)
            f = self.fp12_mul(f, line)
            line, T = self._line(T, self.g2_neg(Q2), P)
            f = self.fp12_mul(f, line)
            return f
        else:
            x = 0xD201000000010000
            for bit in bin(x)[3:]:
                line, T = self._line(T, T, P)
                f = self.fp12_mul(self.fp12_mul(f, f), line)
                if bit == "1":
                    line, T = self._line(T, Q, P)
                    f = self.fp12_mul(f, line)
            return self.fp12_conj(f)  # x < 0

    def final_exp(self, f):
        return self.fp12_pow(f, (self.p ** 12 - 1) // self.r)

    def pair(self, P, Q):
        return self.final_exp(self.miller(P, Q))

    def pairing_product(self, Ps, Qs):
        """curves/curve.go:125-170: returns None when the lengths differ.  The reference
        multiplies n+1 *full* pairings; multiplying the Miller values and exponentiating once
        gives the identical GT element (final_exp is a homomorphism)."""
        if len(Ps) != len(Qs):
            return None
        f = self.fp12_one
        for P, Q in zip(Ps, Qs):
            f = self.fp12_mul(f, self.miller(P, Q))
        return self.final_exp(f)

    # ---- byte layouts ----------------------------------------------------------------
    def _be(self, v):
        return int(v).to_bytes(self.nbytes, "big")

    def marshal_g1(self, P):
        """x||y big-endian (curves/altbn128.go:149-155; bls12 .dat vectors); infinity = zeros."""
        if P is None:
            return bytes(2 * self.nbytes)
        return self._be(P[0]) + self._be(P[1])

    def marshal_g2(self, Q):
        """x_im||x_re||y_im||y_re (curves/altbn128.go:157-158,253-262; bls12_381.go:147-158)."""
        if Q is None:
            return bytes(4 * self.nbytes)
        return self._be(Q[0][1]) + self._be(Q[0][0]) + self._be(Q[1][1]) + self._be(Q[1][0])

    def marshal_gt(self, f):
        """12 Fp values, order of w-powers 5,3,1,4,2,0, each (im, re): the cloudflare bn256
        gfP12 layout (x*omega + y; gfP6 = x tau^2 + y tau + z; gfP2 = x i + y) with tau = w^2
        [upstream layout restated from memory, see SURVEY.md 8c]."""
        out = b""
        for k in (5, 3, 1, 4, 2, 0):
            out += self._be(f[k][1]) + self._be(f[k][0])
        return out

    def unmarshal_g1(self, data):
        n = self.nbytes
        assert len(data) == 2 * n
        if self.name == "bls12" and data[0] & 0x40:
            return None
        x, y = int.from_bytes(data[:n], "big"), int.from_bytes(data[n:], "big")
        return None if (x == 0 and y == 0) else (x, y)

    def unmarshal_g2(self, data):
        n = self.nbytes
        assert len(data) == 4 * n
        if self.name == "bls12" and data[0] & 0x40:
            return None
        v = [int.from_bytes(data[i * n:(i + 1) * n], "big") for i in range(4)]
        return None if not any(v) else ((v[1], v[0]), (v[3], v[2]))

    def unmarshal_gt(self, data):
        n = self.nbytes
        assert len(data) == 12 * n
        v = [int.from_bytes(data[i * n:(i + 1) * n], "big") for i in range(12)]
        f = [None] * 6
        for idx, k in enumerate((5, 3, 1, 4, 2, 0)):
            f[k] = (v[2 * idx + 1], v[2 * idx])
        return f


    # ---- compressed wire formats -----------------------------------------------------
    # altbn128: the reference's own codec (curves/altbn128.go:81-89, 203-221, 296-376).
    #   G1: x (32 B, big-endian), bit 7 of byte 0 set iff 2y > q.          infinity = zeros
    #   G2: x_im || x_re (64 B), bit 7 of x_im[0] iff 2 y_im > q, bit 7 of x_re[0] iff 2 y_re > q.
    # bls12-381: the reference hands (un)marshalling to dis2/bls12 (curves/bls12_381.go:57-63,118-124,242-264), which is
    # not in /root/reference; restated here as the zcash bls12-381 serialisation that library implements
    # [upstream, from memory -- PARITY UNPINNED for this curve]:
    #   G1: x (48 B); byte 0: 0x80 = compressed, 0x40 = infinity, 0x20 = y is the lexicographically larger root.
    #   G2: x_c1 || x_c0 (96 B), same flags; "larger" compares (y_c1, y_c0).
    def sqrt_fp2(self, a):
        """calcComplexQuadRes (curves/hash.go:196-223), Adj & Rodriguez-Henriquez complex method; a = (re, im).
        Returns some square root when a is a square, an arbitrary value otherwise (callers check)."""
        p = self.p
        re, im = a[0] % p, a[1] % p
        if im == 0:
            # the reference only handles a.re being a residue here; the other real case is i*sqrt(-re)
            if self.is_quad_res(re):
                return (self.sqrt_fp(re), 0)
            return (0, self.sqrt_fp((-re) % p))
        lam = self.sqrt_fp((re * re + im * im) % p)
        inv2 = pow(2, -1, p)
        delta = (re + lam) * inv2 % p
        if not self.is_quad_res(delta):
            delta = (re - lam) * inv2 % p
        x0 = self.sqrt_fp(delta)
        if x0 == 0:
            return (0, 0)
        x1 = pow(x0, -1, p) * inv2 % p * im % p
        return (x0, x1)

    def compress_g1(self, P):
        n = self.nbytes
        if self.name == "altbn128":
            if P is None:
                return bytes(n)
            out = bytearray(self._be(P[0]))
            if 2 * P[1] > self.p:
                out[0] |= 0x80
            return bytes(out)
        if P is None:
            return bytes([0xC0]) + bytes(n - 1)
        out = bytearray(self._be(P[0]))
        out[0] |= 0x80 | (0x20 if P[1] > (self.p - 1) // 2 else 0)
        return bytes(out)

    def compress_g2(self, Q):
        n = self.nbytes
        if self.name == "altbn128":
            if Q is None:
                return bytes(2 * n)
            (xr, xi), (yr, yi) = Q
            a, b = bytearray(self._be(xi)), bytearray(self._be(xr))
            if 2 * yi > self.p:
                a[0] |= 0x80
            if 2 * yr > self.p:
                b[0] |= 0x80
            return bytes(a + b)
        if Q is None:
            return bytes([0xC0]) + bytes(2 * n - 1)
        (x0, x1), (y0, y1) = Q
        out = bytearray(self._be(x1) + self._be(x0))
        half = (self.p - 1) // 2
        larger = y1 > half if y1 != 0 else y0 > half
        out[0] |= 0x80 | (0x20 if larger else 0)
        return bytes(out)

    def decompress_g1(self, data):
        """-> (point or None for infinity, ok).  ok is False when x is not the abscissa of a curve point
        (the reference's MakeG1Point(.., true) / upstream Unmarshal rejects it) or the encoding is malformed."""
        n, p = self.nbytes, self.p
        if len(data) != n:
            return None, False
        d = bytearray(data)
        if self.name == "altbn128":
            sgn = d[0] >= 128
            d[0] &= 0x7F
            x = int.from_bytes(d, "big")
            if x == 0:
                return None, True
            if x >= p:
                return None, False
            y2 = self.x_to_y2(x)
            y = self.sqrt_fp(y2)
            if y * y % p != y2:
                return None, False
            if sgn != (2 * y > p):
                y = (p - y) % p
            return (x, y), True
        if not d[0] & 0x80:
            return None, False
        inf, larger = bool(d[0] & 0x40), bool(d[0] & 0x20)
        d[0] &= 0x1F
        x = int.from_bytes(d, "big")
        if inf:
            return None, (x == 0 and not larger)
        if x >= p:
            return None, False
        y2 = self.x_to_y2(x)
        y = self.sqrt_fp(y2)
        if y * y % p != y2:
            return None, False
        if larger != (y > (p - 1) // 2):
            y = (p - y) % p
        return (x, y), True

    def decompress_g2(self, data):
        n, p, F = self.nbytes, self.p, self.F2
        if len(data) != 2 * n:
            return None, False
        d = bytearray(data)
        if self.name == "altbn128":
            si, sr = d[0] >= 128, d[n] >= 128
            d[0] &= 0x7F
            d[n] &= 0x7F
            xi, xr = int.from_bytes(d[:n], "big"), int.from_bytes(d[n:], "big")
            if xi == 0 and xr == 0:
                return None, True
            if xi >= p or xr >= p:
                return None, False
            x = (xr, xi)
            y2 = F.add(F.mul(x, F.mul(x, x)), self.b2)
            yr, yi = self.sqrt_fp2(y2)
            # the reference fixes the two components independently (curves/altbn128.go:355-370)
            if si != (2 * yi > p):
                yi = (p - yi) % p
            if sr != (2 * yr > p):
                yr = (p - yr) % p
            Q = (x, (yr, yi))
            return (Q, True) if self.g2_on_curve(Q) else (None, False)
        if not d[0] & 0x80:
            return None, False
        inf, larger = bool(d[0] & 0x40), bool(d[0] & 0x20)
        d[0] &= 0x1F
        x1, x0 = int.from_bytes(d[:n], "big"), int.from_bytes(d[n:], "big")
        if inf:
            return None, (x0 == 0 and x1 == 0 and not larger)
        if x0 >= p or x1 >= p:
            return None, False
        x = (x0, x1)
        y2 = F.add(F.mul(x, F.mul(x, x)), self.b2)
        y = self.sqrt_fp2(y2)
        if F.mul(y, y) != (y2[0] % p, y2[1] % p):
            return None, False
        half = (p - 1) // 2
        is_larger = y[1] > half if y[1] != 0 else y[0] > half
        if larger != is_larger:
            y = ((p - y[0]) % p, (p - y[1]) % p)
        return (x, y), True

    def in_subgroup_g1(self, P):
        return P is None or self.g1_mul(P, self.r) is None

    def in_subgroup_g2(self, Q):
        return Q is None or self.g2_mul(Q, self.r) is None

    # ---- hash to G1 ------------------------------------------------------------------
    def sqrt_fp(self, a):
        """calcQuadRes, curves/hash.go:178-190 (q = 3 mod 4 only)."""
        return pow(a, (self.p + 1) // 4, self.p)

    def is_quad_res(self, a):
        """curves/hash.go:254-265 (0 counts as a residue)."""
        a %= self.p
        return a == 0 or pow(a, (self.p - 1) // 2, self.p) == 1

    def x_to_y2(self, x):
        return (x * x * x + self.b) % self.p

    def hash_to_g1(self, msg: bytes):
        if self.name == "altbn128":
            return self._try_and_increment_evm(msg)
        return self._hash_bls12(msg)

    def _try_and_increment_evm(self, msg):
        """curves/hash.go:53-77 with EthereumSum256 (curves/altbn128.go:509-522)."""
        q = self.p
        counter = 0
        while True:
            h = keccak256(bytes([counter]) + msg)
            counter = (counter + 1) & 0xFF
            px = int.from_bytes(h[:32], "big") % q
            y2 = self.x_to_y2(px)
            root = self.sqrt_fp(y2)
            if root * root % q == y2:
                py = root
                sign_y = keccak256(bytes([255]) + msg)[31] % 2
                if sign_y == 1:
                    py = q - py
                return (px, py)

    def _sw(self, t):
        """Shallue-van de Woestijne encoding, curves/hash.go:91-167 (non-blind branch)."""
        q, b = self.p, self.b
        root_neg3, z = self.ft_params
        w = pow((t * t + 1 + b) % q, -1, q) * t % q * root_neg3 % q
        x0 = (z - t * w) % q
        x1 = (-x0 - 1) % q
        if self.is_quad_res(self.x_to_y2(x0)):
            x = x0
        elif self.is_quad_res(self.x_to_y2(x1)):
            x = x1
        else:
            x = (pow(w * w % q, -1, q) + 1) % q
        y = self.sqrt_fp(self.x_to_y2(x))
        if self._parity(y) != self._parity(t):
            y = q - y
        return (x, y)

    def _parity(self, x):
        """curves/hash.go:169-172."""
        return x > (self.p - x)

    def _fouque_tibouchi(self, t_bytes):
        """bls12FouqueTibouchi, curves/bls12_381.go:378-393 + fouqueTibouchiG1 hash.go:79-86."""
        t = int.from_bytes(t_bytes, "big") % self.p
        if t == 0:
            return None
        if t == self.ft_root1:
            return self.g1
        if t == self.ft_root2:
            return self.g1_neg(self.g1)
        return self.g1_mul(self._sw(t), self.cofactor)

    def _hash_bls12(self, msg):
        """hashToG1BlindingAbstracted, curves/bls12_381.go:362-376."""
        t1 = hashlib.blake2b(msg + b"G1_0", digest_size=64).digest()
        t2 = hashlib.blake2b(msg + b"G1_1", digest_size=64).digest()
        return self.g1_add(self._fouque_tibouchi(t1), self._fouque_tibouchi(t2))


ALTBN128 = Curve(
    name="altbn128",
    p=21888242871839275222246405745257275088696311157297823662689037894645226208583,
    r=21888242871839275222246405745257275088548364400416034343698204186575808495617,
    b=3,
    xi=(9, 1),
    twist="D",
    g1=(1, 2),
    # curves/altbn128_test.go:26-35 (xi, xr, yi, yr)
    g2=((10857046999023057135944570762232829481370756359578518086990519993285655852781,
         11559732032986387107991004021392285783925812861821192530917403151452391805634),
        (8495653923123431417604973247489272438418190587263600148770280649306958101930,
         4082367875863433681332203403145435568316851327593401208105741076214120093531)),
    nbytes=32,
    cofactor=1,
)
ALTBN128.ft_params = (4407920970296243842837207485651524041948558517760411303933,
                      2203960485148121921418603742825762020974279258880205651966)

BLS12_381 = Curve(
    name="bls12",
    p=0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab,
    r=52435875175126190479447740508185965837690552500527637822603658699938581184513,
    b=4,
    xi=(1, 1),
    twist="M",
    g1=(0x17f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac586c55e83ff97a1aeffb3af00adb22c6bb,
        0x08b3f481e3aaa0f1a09e30ed741d8ae4fcf5e095d5d00af600db18cb2c04b3edd03cc744a2888ae40caa232946c5e7e1),
    g2=((0x024aa2b2f08f0a91260805272dc51051c6e47ad4fa403b02b4510b647ae3d1770bac0326a805bbefd48056c8c121bdb8,
         0x13e02b6052719f607dacd3a088274f65596bd0d09920b61ab5da61bbdc7f5049334cf11213945d57e5ac7d055d042b7e),
        (0x0ce5d527727d6e118cc9cdc6da2e351aadfd9baa8cbdd3a76d429a695160d12c923ac9cc3baca289e193548608b82801,
         0x0606c4a02ea734cc32acd2b02bc28b99cb3e287e85a763af267492ab572e99ab3f370d275cec1da1aaa9075ff05f79be)),
    nbytes=48,
    cofactor=76329603384216526031706109802092473003,
)
# curves/bls12_381.go:333-346
BLS12_381.ft_params = (
    1586958781458431025242759403266842894121773480562120986020912974854563298150952611241517463240701,
    793479390729215512621379701633421447060886740281060493010456487427281649075476305620758731620350)
BLS12_381.ft_root1 = 248294325734266649657405162895821171812231848760181225578082735178502750823719347628762635478508544819911854747095
BLS12_381.ft_root2 = 3754115229487400743760384662840082984744650971178826659753975400945528899667118516813924993650507119217982417812692

CURVES = {"altbn128": ALTBN128, "bls12": BLS12_381}


# --------------------------------------------------------------------------------------
# curves/curve.go package-level functions and the bgls scheme layer
# --------------------------------------------------------------------------------------


def aggregate_points(curve: Curve, pts, group="g2"):
    """AggregatePoints, curves/curve.go:73-110: pairwise tree.  len 1 returns the point,
    len 0 never terminates in the reference (the loop at :94-108 cannot reach length 1)."""
    if len(pts) == 0:
        raise ValueError("reference AggregatePoints([]) does not terminate (curve.go:94-108)")
    add = curve.g1_add if group == "g1" else curve.g2_add
    level = list(pts)
    if len(level) == 2:
        return add(level[0], level[1])
    while True:
        nxt = []
        for i in range(0, len(level), 2):
            nxt.append(level[i] if i + 1 >= len(level) else add(level[i], level[i + 1]))
        level = nxt
        if len(level) == 1:
            return level[0]


def keygen(curve: Curve, rng):
    """bgls/bgls.go:30-43 (rng replaces crypto/rand so tests are reproducible)."""
    x = rng.randrange(curve.r)
    return x, curve.g2_mul(curve.g2, x)


def sign(curve: Curve, sk, msg):
    """bgls/bgls.go:46-56."""
    return curve.g1_mul(curve.hash_to_g1(msg), sk)


def is_gt_identity(curve: Curve, f):
    return f == curve.fp12_one


def verify_single(curve: Curve, sig, pubkey, msg):
    """VerifySingleSignatureCustHash, bgls/bgls.go:65-70."""
    h = curve.g1_neg(curve.hash_to_g1(msg))
    paired = curve.pairing_product([h, sig], [pubkey, curve.g2])
    return is_gt_identity(curve, paired)


def verify_agg_sig(curve: Curve, aggsig, keys, msgs, allow_duplicates=False):
    """verifyAggSig, bgls/bgls.go:94-119."""
    if len(keys) != len(msgs):
        return False
    if not allow_duplicates and len(set(bytes(m) for m in msgs)) != len(msgs):
        return False
    pts1 = [curve.hash_to_g1(m) for m in msgs] + [curve.g1_neg(aggsig)]
    pts2 = list(keys) + [curve.g2]
    paired = curve.pairing_product(pts1, pts2)
    return paired is not None and is_gt_identity(curve, paired)


def verify_multi_sig(curve: Curve, aggsig, keys, msg):
    """verifyMultiSignature, bgls/bgls.go:89-92."""
    return verify_single(curve, aggsig, aggregate_points(curve, keys, "g2"), msg)


def kosk_verify_multi_sig(curve: Curve, aggsig, keys, msg):
    """KoskVerifyMultiSignature, bgls/blsKosk.go:117-120."""
    return verify_multi_sig(curve, aggsig, keys, b"\x01" + msg)
