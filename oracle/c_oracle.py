"""TEST INFRASTRUCTURE (oracle) -- ctypes loader for oracle/liboracle.so (the C restatement).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use this.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

ALTBN128, BLS12 = 0, 1
FP_BYTES = {ALTBN128: 32, BLS12: 48}


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE, "liboracle.so"])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        L = ctypes.CDLL(path)
        u8p, sz, i = ctypes.c_char_p, ctypes.c_size_t, ctypes.c_int
        L.orc_pairing_product.argtypes = [i, u8p, u8p, sz, u8p, i, i]
        L.orc_miller_product.argtypes = [i, u8p, u8p, sz, u8p, i]
        L.orc_fp12_product.argtypes = [i, u8p, sz, i, u8p]
        L.orc_fp12_pow.argtypes = [i, u8p, ctypes.POINTER(ctypes.c_uint64), i, u8p]
        L.orc_aggregate.argtypes = [i, i, u8p, sz, u8p, i]
        L.orc_scale_points.argtypes = [i, i, u8p, u8p, sz, u8p, i]
        L.orc_on_curve.argtypes = [i, i, u8p]
        L.orc_fpmul_count.restype = ctypes.c_uint64
        _LIB = L
    return _LIB


def _buf(n):
    return ctypes.create_string_buffer(n)


def pairing_product(curve, g1: bytes, g2: bytes, n: int, nthreads=1, mode=0) -> bytes:
    out = _buf(12 * FP_BYTES[curve])
    rc = lib().orc_pairing_product(curve, g1, g2, n, out, nthreads, mode)
    assert rc == 0
    return out.raw


def miller_product(curve, g1: bytes, g2: bytes, n: int, nthreads=1) -> bytes:
    out = _buf(12 * FP_BYTES[curve])
    assert lib().orc_miller_product(curve, g1, g2, n, out, nthreads) == 0
    return out.raw


def fp12_product(curve, blobs: bytes, k: int, do_final: bool) -> bytes:
    out = _buf(12 * FP_BYTES[curve])
    assert lib().orc_fp12_product(curve, blobs, k, int(do_final), out) == 0
    return out.raw


def fp12_pow(curve, f: bytes, e: int) -> bytes:
    nl = (e.bit_length() + 63) // 64
    arr = (ctypes.c_uint64 * nl)(*[(e >> (64 * j)) & (2**64 - 1) for j in range(nl)])
    out = _buf(12 * FP_BYTES[curve])
    assert lib().orc_fp12_pow(curve, f, arr, nl, out) == 0
    return out.raw


def aggregate(curve, group, pts: bytes, n: int, nthreads=1) -> bytes:
    out = _buf(2 * group * FP_BYTES[curve])
    assert lib().orc_aggregate(curve, group, pts, n, out, nthreads) == 0
    return out.raw


def scale_points(curve, group, pts: bytes, scalars: bytes, n: int, nthreads=1) -> bytes:
    out = _buf(n * 2 * group * FP_BYTES[curve])
    assert lib().orc_scale_points(curve, group, pts, scalars, n, out, nthreads) == 0
    return out.raw


def on_curve(curve, group, pt: bytes) -> bool:
    return bool(lib().orc_on_curve(curve, group, pt))


def fpmul_count() -> int:
    return lib().orc_fpmul_count()


def fpmul_reset():
    lib().orc_fpmul_reset()
