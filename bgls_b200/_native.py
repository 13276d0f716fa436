"""ctypes binding of the C ABI in include/bgls_b200.h (plumbing only: no arithmetic here).

The library is the product; if it is missing or no CUDA device is usable this module raises --
there is deliberately no CPU fallback."""
from __future__ import annotations

import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("BGLS_LIB_PATH") or os.path.join(HERE, "lib", "libbgls_b200.so")   # BGLS_LIB_PATH: A/B builds of the same ABI

ALTBN128, BLS12_381 = 0, 1
G1, G2 = 1, 2
FP_BYTES = {ALTBN128: 32, BLS12_381: 48}

EXPORTS = [
    "bgls_ctx_create", "bgls_ctx_destroy", "bgls_last_error", "bgls_version", "bgls_pairing_product", "bgls_pair",
    "bgls_gt_mul", "bgls_aggregate_points", "bgls_scale_points", "bgls_miller_product", "bgls_final_exp_product",
    "bgls_pairing_check_batch", "bgls_pairing_product_dev", "bgls_miller_product_dev", "bgls_final_exp_product_dev",
    "bgls_aggregate_points_dev", "bgls_scale_points_dev", "bgls_pairing_check_batch_dev", "bgls_launch_count", "bgls_hash_to_g1", "bgls_hash_to_g1_dev", "bgls_set_profiling", "bgls_last_kernel_ms", "bgls_intpipe_peak",
    "bgls_compress_points", "bgls_compress_points_dev", "bgls_decompress_points", "bgls_decompress_points_dev",
    "bgls_verify_aggregate_signature", "bgls_verify_multi_signature",
    "bgls_exchange_create", "bgls_exchange_connect", "bgls_exchange_error", "bgls_miller_product_exchange_dev",
    "bgls_final_exp_exchanged_dev", "bgls_validate_points", "bgls_validate_points_dev", "bgls_gt_pow",
]

_lib = None


class BglsError(RuntimeError):
    pass


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise BglsError(f"{LIB_PATH} not built: run `python -m bgls_b200.build` (no CPU fallback exists)")
    L = ctypes.CDLL(LIB_PATH)
    vp, sz, i, cp = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p
    ip = ctypes.POINTER(ctypes.c_int)
    L.bgls_ctx_create.argtypes = [i, ctypes.POINTER(vp)]
    L.bgls_ctx_destroy.argtypes = [vp]
    L.bgls_ctx_destroy.restype = None
    L.bgls_last_error.argtypes = [vp]
    L.bgls_last_error.restype = ctypes.c_char_p
    L.bgls_version.restype = ctypes.c_char_p
    L.bgls_pairing_product.argtypes = [vp, i, cp, cp, sz, cp, ip]
    L.bgls_pair.argtypes = [vp, i, cp, cp, cp]
    L.bgls_gt_mul.argtypes = [vp, i, cp, cp, cp]
    L.bgls_aggregate_points.argtypes = [vp, i, i, cp, sz, cp]
    L.bgls_scale_points.argtypes = [vp, i, i, cp, cp, sz, cp]
    L.bgls_miller_product.argtypes = [vp, i, cp, cp, sz, cp]
    L.bgls_final_exp_product.argtypes = [vp, i, cp, sz, cp, ip]
    L.bgls_pairing_check_batch.argtypes = [vp, i, cp, cp, ctypes.POINTER(ctypes.c_uint64), sz, cp]
    L.bgls_pairing_product_dev.argtypes = [vp, i, vp, vp, sz, vp, vp, vp]
    L.bgls_miller_product_dev.argtypes = [vp, i, vp, vp, sz, vp, vp]
    L.bgls_final_exp_product_dev.argtypes = [vp, i, vp, sz, vp, vp, vp]
    L.bgls_aggregate_points_dev.argtypes = [vp, i, i, vp, sz, vp, vp]
    L.bgls_scale_points_dev.argtypes = [vp, i, i, vp, vp, sz, vp, vp]
    L.bgls_pairing_check_batch_dev.argtypes = [vp, i, vp, vp, vp, sz, sz, vp, vp]
    L.bgls_hash_to_g1.argtypes = [vp, i, cp, ctypes.POINTER(ctypes.c_uint64), sz, cp]
    L.bgls_hash_to_g1_dev.argtypes = [vp, i, vp, vp, sz, vp, vp]
    L.bgls_compress_points.argtypes = [vp, i, i, cp, sz, cp]
    L.bgls_compress_points_dev.argtypes = [vp, i, i, vp, sz, vp, vp]
    L.bgls_decompress_points.argtypes = [vp, i, i, cp, sz, i, cp, cp]
    L.bgls_decompress_points_dev.argtypes = [vp, i, i, vp, sz, i, vp, vp, vp]
    L.bgls_verify_aggregate_signature.argtypes = [vp, i, cp, ctypes.POINTER(ctypes.c_uint64), sz, cp, cp, i, ip]
    L.bgls_verify_multi_signature.argtypes = [vp, i, cp, sz, cp, sz, cp, ip]
    L.bgls_exchange_create.argtypes = [vp, i, i, i, cp]
    L.bgls_exchange_connect.argtypes = [vp, i, cp]
    L.bgls_exchange_error.argtypes = [vp, ip]
    L.bgls_miller_product_exchange_dev.argtypes = [vp, i, vp, vp, sz, i, ctypes.c_uint64, vp]
    L.bgls_final_exp_exchanged_dev.argtypes = [vp, i, i, ctypes.c_uint64, vp, vp, vp]
    L.bgls_validate_points.argtypes = [vp, i, i, cp, sz, i, cp]
    L.bgls_validate_points_dev.argtypes = [vp, i, i, vp, sz, i, vp, vp]
    L.bgls_gt_pow.argtypes = [vp, i, cp, cp, i, cp]
    L.bgls_launch_count.argtypes = [vp]
    L.bgls_launch_count.restype = ctypes.c_uint64
    L.bgls_set_profiling.argtypes = [vp, i]
    L.bgls_last_kernel_ms.argtypes = [vp, ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_float)]
    L.bgls_intpipe_peak.argtypes = [vp, ctypes.POINTER(ctypes.c_double)]
    _lib = L
    return L


class Context:
    """One engine context bound to one CUDA device (bgls_ctx)."""

    def __init__(self, device: int = 0):
        L = load()
        h = ctypes.c_void_p()
        rc = L.bgls_ctx_create(device, ctypes.byref(h))
        if rc != 0:
            raise BglsError(f"bgls_ctx_create(device={device}) failed with {rc}: no usable CUDA device (no CPU fallback)")
        self._h, self._L, self.device = h, L, device

    def close(self):
        if getattr(self, "_h", None):
            self._L.bgls_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, rc):
        if rc != 0:
            raise BglsError(f"bgls error {rc}: {self._L.bgls_last_error(self._h).decode()}")

    @property
    def launches(self) -> int:
        return int(self._L.bgls_launch_count(self._h))

    def set_profiling(self, on: bool):
        self._chk(self._L.bgls_set_profiling(self._h, int(on)))

    def last_kernel_ms(self):
        a, b = ctypes.c_float(0), ctypes.c_float(0)
        self._chk(self._L.bgls_last_kernel_ms(self._h, ctypes.byref(a), ctypes.byref(b)))
        return a.value, b.value

    def intpipe_peak(self) -> float:
        v = ctypes.c_double(0)
        self._chk(self._L.bgls_intpipe_peak(self._h, ctypes.byref(v)))
        return v.value

    def pairing_product_ptr(self, curve, h_g1: int, h_g2: int, n: int, h_out: int, h_flag: int):
        """Host-buffer call on raw host addresses (e.g. pinned torch tensors)."""
        self._chk(self._L.bgls_pairing_product(self._h, curve, h_g1, h_g2, n, h_out, ctypes.cast(h_flag, ctypes.POINTER(ctypes.c_int))))

    # ---- host-buffer API (bytes in / bytes out)
    def pairing_product(self, curve, g1: bytes, g2: bytes, n: int):
        out = ctypes.create_string_buffer(12 * FP_BYTES[curve])
        flag = ctypes.c_int(0)
        self._chk(self._L.bgls_pairing_product(self._h, curve, g1, g2, n, out, ctypes.byref(flag)))
        return out.raw, bool(flag.value)

    def pair(self, curve, g1: bytes, g2: bytes) -> bytes:
        out = ctypes.create_string_buffer(12 * FP_BYTES[curve])
        self._chk(self._L.bgls_pair(self._h, curve, g1, g2, out))
        return out.raw

    def gt_mul(self, curve, a: bytes, b: bytes) -> bytes:
        out = ctypes.create_string_buffer(12 * FP_BYTES[curve])
        self._chk(self._L.bgls_gt_mul(self._h, curve, a, b, out))
        return out.raw

    def aggregate_points(self, curve, group, pts: bytes, n: int) -> bytes:
        out = ctypes.create_string_buffer(2 * group * FP_BYTES[curve])
        self._chk(self._L.bgls_aggregate_points(self._h, curve, group, pts, n, out))
        return out.raw

    def scale_points(self, curve, group, pts: bytes, scalars: bytes, n: int) -> bytes:
        out = ctypes.create_string_buffer(max(1, n * 2 * group * FP_BYTES[curve]))
        self._chk(self._L.bgls_scale_points(self._h, curve, group, pts, scalars, n, out))
        return out.raw[: n * 2 * group * FP_BYTES[curve]]

    def verify_aggregate_signature(self, curve, msgs, keys: bytes, sig: bytes, allow_duplicates: bool = False) -> bool:
        """verifyAggSig (bgls/bgls.go:94-119) in one engine call: msgs is a list of bytes, keys the packed G2 records."""
        n = len(msgs)
        offs = [0]
        for m in msgs:
            offs.append(offs[-1] + len(m))
        off = (ctypes.c_uint64 * (n + 1))(*offs)
        ok = ctypes.c_int(0)
        self._chk(self._L.bgls_verify_aggregate_signature(self._h, curve, b"".join(msgs), off, n, keys, sig,
                                                          1 if allow_duplicates else 0, ctypes.byref(ok)))
        return bool(ok.value)

    def verify_multi_signature(self, curve, msg: bytes, keys: bytes, n: int, sig: bytes) -> bool:
        """verifyMultiSignature (bgls/bgls.go:89-92) in one engine call."""
        ok = ctypes.c_int(0)
        self._chk(self._L.bgls_verify_multi_signature(self._h, curve, msg, len(msg), keys, n, sig, ctypes.byref(ok)))
        return bool(ok.value)

    def verify_multi_signature_ptr(self, curve, msg: bytes, h_keys: int, n: int, sig: bytes) -> bool:
        """Same call with the keys at a raw host address (pinned buffer of the bench)."""
        ok = ctypes.c_int(0)
        self._chk(self._L.bgls_verify_multi_signature(self._h, curve, msg, len(msg), ctypes.c_char_p(h_keys), n, sig, ctypes.byref(ok)))
        return bool(ok.value)

    def verify_aggregate_signature_ptr(self, curve, h_msgs: int, h_offsets: int, n: int, h_keys: int, h_sig: int, allow_duplicates: bool = False) -> bool:
        """Same call on raw host pointers (pinned buffers of the bench)."""
        ok = ctypes.c_int(0)
        self._chk(self._L.bgls_verify_aggregate_signature(self._h, curve, ctypes.c_char_p(h_msgs), ctypes.cast(h_offsets, ctypes.POINTER(ctypes.c_uint64)),
                                                          n, ctypes.c_char_p(h_keys), ctypes.c_char_p(h_sig), 1 if allow_duplicates else 0, ctypes.byref(ok)))
        return bool(ok.value)

    def validate_points(self, curve, group, pts: bytes, n: int, reference: bool = True) -> list:
        """What the reference checks when it builds a Point (bgls_validate_points): list of n booleans."""
        ok = ctypes.create_string_buffer(max(n, 1))
        self._chk(self._L.bgls_validate_points(self._h, curve, group, pts, n, 1 if reference else 0, ok))
        return [b != 0 for b in ok.raw[:n]]

    def gt_pow(self, curve, a: bytes, e: int) -> bytes:
        """PointT.Mul: a^e in GT."""
        out = ctypes.create_string_buffer(12 * FP_BYTES[curve])
        self._chk(self._L.bgls_gt_pow(self._h, curve, a, abs(e).to_bytes(32, "big"), 1 if e < 0 else 0, out))
        return out.raw

    def compress_points(self, curve, group, pts: bytes, n: int) -> bytes:
        """n uncompressed records -> n compressed records (Point.Marshal)."""
        out = ctypes.create_string_buffer(max(1, n * group * FP_BYTES[curve]))
        self._chk(self._L.bgls_compress_points(self._h, curve, group, pts, n, out))
        return out.raw[: n * group * FP_BYTES[curve]]

    def decompress_points(self, curve, group, data: bytes, n: int, check_subgroup: bool = False):
        """n compressed records -> (n uncompressed records, list of ok flags) (UnmarshalG1 / UnmarshalG2)."""
        out = ctypes.create_string_buffer(max(1, n * 2 * group * FP_BYTES[curve]))
        ok = ctypes.create_string_buffer(max(1, n))
        self._chk(self._L.bgls_decompress_points(self._h, curve, group, data, n, 1 if check_subgroup else 0, out, ok))
        return out.raw[: n * 2 * group * FP_BYTES[curve]], [bool(b) for b in ok.raw[:n]]

    def miller_product(self, curve, g1: bytes, g2: bytes, n: int) -> bytes:
        out = ctypes.create_string_buffer(12 * FP_BYTES[curve])
        self._chk(self._L.bgls_miller_product(self._h, curve, g1, g2, n, out))
        return out.raw

    def final_exp_product(self, curve, partials: bytes, k: int):
        out = ctypes.create_string_buffer(12 * FP_BYTES[curve])
        flag = ctypes.c_int(0)
        self._chk(self._L.bgls_final_exp_product(self._h, curve, partials, k, out, ctypes.byref(flag)))
        return out.raw, bool(flag.value)

    def hash_to_g1(self, curve, msgs) -> bytes:
        """HashToG1 of every message (list of bytes) -> concatenated uncompressed G1 records."""
        n = len(msgs)
        offs = [0]
        for m in msgs:
            offs.append(offs[-1] + len(m))
        off = (ctypes.c_uint64 * (n + 1))(*offs)
        blob = b"".join(bytes(m) for m in msgs) or b"\0"
        out = ctypes.create_string_buffer(max(1, n * 2 * FP_BYTES[curve]))
        self._chk(self._L.bgls_hash_to_g1(self._h, curve, blob, off, n, out))
        return out.raw[: n * 2 * FP_BYTES[curve]]

    def pairing_check_batch(self, curve, g1: bytes, g2: bytes, offsets) -> list:
        nb = len(offsets) - 1
        off = (ctypes.c_uint64 * (nb + 1))(*offsets)
        out = ctypes.create_string_buffer(max(1, nb))
        self._chk(self._L.bgls_pairing_check_batch(self._h, curve, g1, g2, off, nb, out))
        return [bool(b) for b in out.raw[:nb]]

    # ---- device-resident API (raw device pointers + cudaStream_t handles as ints)
    # ---- peer-memory exchange (multi-GPU)
    def exchange_create(self, world: int, rank: int, lanes: int) -> bytes:
        h = ctypes.create_string_buffer(64)
        self._chk(self._L.bgls_exchange_create(self._h, world, rank, lanes, h))
        return h.raw

    def exchange_connect(self, peer_rank: int, handle: bytes):
        self._chk(self._L.bgls_exchange_connect(self._h, peer_rank, handle))

    def exchange_error(self) -> int:
        e = ctypes.c_int(0)
        self._chk(self._L.bgls_exchange_error(self._h, ctypes.byref(e)))
        return e.value

    def miller_product_exchange_dev(self, curve, d_g1, d_g2, n, lane, epoch, stream):
        self._chk(self._L.bgls_miller_product_exchange_dev(self._h, curve, d_g1, d_g2, n, lane, epoch, stream))

    def final_exp_exchanged_dev(self, curve, lane, epoch, d_out, d_flag, stream):
        self._chk(self._L.bgls_final_exp_exchanged_dev(self._h, curve, lane, epoch, d_out, d_flag, stream))

    def pairing_product_dev(self, curve, d_g1, d_g2, n, d_out, d_flag, stream):
        self._chk(self._L.bgls_pairing_product_dev(self._h, curve, d_g1, d_g2, n, d_out, d_flag, stream))

    def miller_product_dev(self, curve, d_g1, d_g2, n, d_out, stream):
        self._chk(self._L.bgls_miller_product_dev(self._h, curve, d_g1, d_g2, n, d_out, stream))

    def final_exp_product_dev(self, curve, d_partials, k, d_out, d_flag, stream):
        self._chk(self._L.bgls_final_exp_product_dev(self._h, curve, d_partials, k, d_out, d_flag, stream))

    def aggregate_points_dev(self, curve, group, d_pts, n, d_out, stream):
        self._chk(self._L.bgls_aggregate_points_dev(self._h, curve, group, d_pts, n, d_out, stream))

    def scale_points_dev(self, curve, group, d_pts, d_sc, n, d_out, stream):
        self._chk(self._L.bgls_scale_points_dev(self._h, curve, group, d_pts, d_sc, n, d_out, stream))

    def pairing_check_batch_dev(self, curve, d_g1, d_g2, d_off, nbatch, total, d_ok, stream):
        self._chk(self._L.bgls_pairing_check_batch_dev(self._h, curve, d_g1, d_g2, d_off, nbatch, total, d_ok, stream))
