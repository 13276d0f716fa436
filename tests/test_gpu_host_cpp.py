"""The C++ host mirror of the reference interface (bgls_b200/host/bgls.hpp) on a GPU: the reference's own scheme and
curve tests rendered in C++ (tests/host_cpp/test_bgls_host.cpp), and a deterministic transcript whose every line is
recomputed here with the oracle -- byte parity of keys, signatures, aggregates and GT values through the C++ API."""
import os
import subprocess

import pytest

from oracle import bgls_oracle as O
from oracle import c_oracle as C
from parity_util import CURVES, scalars_bytes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "host_cpp", "test_bgls_host")


def build_host_test():
    from bgls_b200 import build
    build.build(verbose=False)
    src = EXE + ".cpp"
    hdr = os.path.join(ROOT, "bgls_b200", "host", "bgls.hpp")
    if not os.path.exists(EXE) or max(os.path.getmtime(src), os.path.getmtime(hdr)) > os.path.getmtime(EXE):
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-Wall", "-o", EXE, src, "-L" + os.path.join(ROOT, "bgls_b200", "lib"),
                               "-lbgls_b200", "-Wl,-rpath,$ORIGIN/../../bgls_b200/lib"])
    return EXE


def test_host_cpp_builds_and_refuses_without_gpu():
    """CPU: the mirror compiles against include/bgls_b200.h and links the library; without a device it must fail
    loudly (no CPU fallback)."""
    import torch
    exe = build_host_test()
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu tests")
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 2 and "no CUDA device" in r.stderr


def test_host_cpp_blake2xb_matches_python():
    """CPU: the C++ BLAKE2Xb of the hashed aggregation exponents against bgls_b200/blake2x.py (itself pinned on hashlib)."""
    from bgls_b200.blake2x import blake2xb
    r = subprocess.run([build_host_test(), "--blake2x"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = r.stdout.strip().splitlines()
    assert len(lines) == 24
    for line in lines:
        n, length, hx = line.split()
        n, length = int(n), int(length)
        assert hx == blake2xb(bytes((7 * k + n) & 0xFF for k in range(n)), length).hex(), (n, length)


@pytest.mark.gpu
def test_reference_tests_in_cpp():
    r = subprocess.run([build_host_test()], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    assert r.stdout.strip() == "ok"


@pytest.mark.gpu
def test_transcript_matches_oracle():
    r = subprocess.run([build_host_test(), "--dump"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    got = dict(line.split(" ", 1) for line in r.stdout.strip().splitlines())
    for cid, c in CURVES:
        name = "altbn128" if cid == 0 else "bls12"
        n = 4
        sks = [1000003 * (i + 1) + 7 for i in range(n)]
        msgs = [b"msg%d" % i for i in range(n)]
        g2 = c.marshal_g2(c.g2)
        hs = [c.marshal_g1(c.hash_to_g1(m)) for m in msgs]
        pks = [C.scale_points(cid, 2, g2, scalars_bytes([sk]), 1) for sk in sks]
        sigs = [C.scale_points(cid, 1, h, scalars_bytes([sk]), 1) for h, sk in zip(hs, sks)]
        for i in range(n):
            assert got[f"{name}.pk{i}"] == pks[i].hex(), (name, "pk", i)
            assert got[f"{name}.sig{i}"] == sigs[i].hex(), (name, "sig", i)
        assert got[f"{name}.aggsig"] == C.aggregate(cid, 1, b"".join(sigs), n, 1).hex()
        assert got[f"{name}.aggkey"] == C.aggregate(cid, 2, b"".join(pks), n, 1).hex()
        assert got[f"{name}.product"] == C.pairing_product(cid, b"".join(hs), b"".join(pks), n, 1, 0).hex()
        assert got[f"{name}.pair"] == C.pairing_product(cid, hs[0], pks[0], 1, 1, 0).hex()
        # hashed aggregation exponents: the C++ BLAKE2Xb against the Python one, the scaled sum against the oracle
        from bgls_b200.blake2x import blake2xb
        ex = blake2xb(b"".join(pks), 16 * n)
        assert got[f"{name}.hae_exponents"] == ex.hex()
        scaled = C.scale_points(cid, 1, b"".join(sigs), b"".join(bytes(16) + ex[16 * i:16 * i + 16] for i in range(n)), n, 1)
        assert got[f"{name}.hae_aggsig"] == C.aggregate(cid, 1, scaled, n, 1).hex()
        assert got[f"{name}.verify"] == "1" and got[f"{name}.verify_bad"] == "0"
