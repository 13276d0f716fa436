"""BLAKE2Xb extendable-output function, as golang.org/x/crypto/blake2b.NewXOF(size, nil) computes it -- the hash
G^n -> R^n of the reference's hashed aggregation exponents (bgls/blsHAE.go:81-93).  Host glue only (the exponents are
a few bytes per key; the scalings and pairings they feed run on the GPU).

BLAKE2X (Aumasson, Neves, Wilcox-O'Hearn, Winnerlein 2016), restated from the specification:
  H0    = BLAKE2b(M)  with digest_length 64 and the XOF length in bytes 12..15 of the parameter block;
  out_i = BLAKE2b(H0) with digest_length min(64, remaining), fanout 0, depth 0, leaf_length 64, node_offset i,
          XOF length, node_depth 0, inner_length 64.
hashlib.blake2b computes H0 (its 64-bit node_offset covers the XOF-length field) but rejects depth 0, so the output
blocks use the one-block BLAKE2b compression below; `_selfcheck` pins that compression against hashlib.
Parity with the Go library is *unpinned*: the reference holds no HAE known-answer vector and Go is not available."""
from __future__ import annotations

import hashlib
import struct

_IV = (0x6A09E667F3BCC908, 0xBB67AE8584CAA73B, 0x3C6EF372FE94F82B, 0xA54FF53A5F1D36F1,
       0x510E527FADE682D1, 0x9B05688C2B3E6C1F, 0x1F83D9ABFB41BD6B, 0x5BE0CD19137E2179)
_SIGMA = (
    (0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15), (14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3),
    (11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4), (7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8),
    (9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13), (2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9),
    (12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11), (13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10),
    (6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5), (10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0),
    (0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15), (14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3),
)
_M = (1 << 64) - 1


def _rotr(x, n):
    return ((x >> n) | (x << (64 - n))) & _M


def _compress(h, block, t, last):
    m = struct.unpack("<16Q", block)
    v = list(h) + list(_IV)
    v[12] ^= t & _M
    v[13] ^= t >> 64
    if last:
        v[14] ^= _M

    def g(a, b, c, d, x, y):
        v[a] = (v[a] + v[b] + x) & _M
        v[d] = _rotr(v[d] ^ v[a], 32)
        v[c] = (v[c] + v[d]) & _M
        v[b] = _rotr(v[b] ^ v[c], 24)
        v[a] = (v[a] + v[b] + y) & _M
        v[d] = _rotr(v[d] ^ v[a], 16)
        v[c] = (v[c] + v[d]) & _M
        v[b] = _rotr(v[b] ^ v[c], 63)
    for r in range(12):
        s = _SIGMA[r]
        g(0, 4, 8, 12, m[s[0]], m[s[1]])
        g(1, 5, 9, 13, m[s[2]], m[s[3]])
        g(2, 6, 10, 14, m[s[4]], m[s[5]])
        g(3, 7, 11, 15, m[s[6]], m[s[7]])
        g(0, 5, 10, 15, m[s[8]], m[s[9]])
        g(1, 6, 11, 12, m[s[10]], m[s[11]])
        g(2, 7, 8, 13, m[s[12]], m[s[13]])
        g(3, 4, 9, 14, m[s[14]], m[s[15]])
    return [h[i] ^ v[i] ^ v[i + 8] for i in range(8)]


def _blake2b_params(data: bytes, params: bytes) -> bytes:
    """Unkeyed BLAKE2b of `data` under an explicit 64-byte parameter block; returns the full 64-byte state."""
    p = struct.unpack("<8Q", params)
    h = [iv ^ x for iv, x in zip(_IV, p)]
    n = max(1, -(-len(data) // 128))
    for i in range(n):
        chunk = data[128 * i:128 * (i + 1)]
        last = i == n - 1
        t = len(data) if last else 128 * (i + 1)
        h = _compress(h, chunk.ljust(128, b"\0"), t, last)
    return struct.pack("<8Q", *h)


def _param_block(digest_len, fanout, depth, leaf_len, node_offset, xof_len, node_depth, inner_len) -> bytes:
    return struct.pack("<BBBBIIIBB", digest_len, 0, fanout, depth, leaf_len, node_offset, xof_len, node_depth, inner_len) + bytes(46)


def blake2xb(data: bytes, out_len: int) -> bytes:
    """out_len bytes of BLAKE2Xb(data), unkeyed; out_len < 2^32 - 1 (the known-length mode NewXOF(size) uses)."""
    assert 0 < out_len < 0xFFFFFFFF
    h0 = hashlib.blake2b(data, digest_size=64, node_offset=out_len << 32).digest()
    out, i = [], 0
    while 64 * i < out_len:
        j = min(64, out_len - 64 * i)
        out.append(_blake2b_params(h0, _param_block(j, 0, 0, 64, i, out_len, 0, 64))[:j])
        i += 1
    return b"".join(out)


def _selfcheck():
    """The compression / parameter-block code above against hashlib on parameter sets hashlib accepts."""
    for n in (0, 1, 64, 127, 128, 129, 300):
        d = bytes((7 * k + n) & 0xFF for k in range(n))
        for size in (16, 32, 64):
            assert _blake2b_params(d, _param_block(size, 1, 1, 0, 0, 0, 0, 0))[:size] == hashlib.blake2b(d, digest_size=size).digest()
        assert _blake2b_params(d, _param_block(64, 1, 1, 0, 0, 99, 0, 0)) == hashlib.blake2b(d, digest_size=64, node_offset=99 << 32).digest()
        assert (_blake2b_params(d, _param_block(48, 2, 3, 64, 5, 77, 1, 64))[:48] ==
                hashlib.blake2b(d, digest_size=48, fanout=2, depth=3, leaf_size=64, node_offset=5 | (77 << 32), node_depth=1, inner_size=64).digest())
    return True
