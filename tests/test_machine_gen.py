"""tools/gen_machine.py: the generated programs simulated with exact integers against the oracle, the
static worst-case bound verification, and the committed tables being up to date."""
import os
import random
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import gen_machine as GM  # noqa: E402
from oracle import bgls_oracle as O  # noqa: E402

CASES = [(GM.BN, O.ALTBN128), (GM.BLS, O.BLS12_381)]


@pytest.fixture(scope="module")
def built():
    return {cfg.name: GM.build_all(cfg) for cfg, _ in CASES}


@pytest.mark.parametrize("cfg,c", CASES)
def test_static_bounds(built, cfg, c):
    gens, io = built[cfg.name]
    bits = GM.verify_all(cfg, gens, io)
    assert max(bits.values()) < cfg.W * cfg.L


@pytest.mark.parametrize("cfg,c", CASES)
def test_programs_against_oracle(built, cfg, c):
    gens, io = built[cfg.name]
    gm, gf = gens["M"], gens["F"]
    rng = random.Random(5)
    P = c.g1_mul(c.g1, rng.randrange(c.r))
    Q = c.g2_mul(c.g2, rng.randrange(c.r))
    sim = GM.Sim(gm)
    raw = io["miller_in"]
    for k, v in (("xP", P[0]), ("yP", P[1]), ("xQ.x", Q[0][0]), ("xQ.y", Q[0][1]), ("yQ.x", Q[1][0]), ("yQ.y", Q[1][1])):
        sim.set(raw[k], v)
    sim.run("MILLER")
    Rinv = pow(cfg.R, -1, cfg.p)
    FA = io["FA"]
    f = [(sim.get(FA[k][0]) * Rinv % cfg.p, sim.get(FA[k][1]) * Rinv % cfg.p) for k in range(6)]
    want = c.pair(P, Q)
    assert c.final_exp(f) == want
    # the pipelined 32-lane program (slot file P, signed terms) computes the same pairing
    simp = GM.Sim(gens["P"])
    for k, v in (("xP", P[0]), ("yP", P[1]), ("xQ.x", Q[0][0]), ("xQ.y", Q[0][1]), ("yQ.x", Q[1][0]), ("yQ.y", Q[1][1])):
        simp.set(io["p_miller_in"][k], v)
    simp.run("MILLER")
    PFA = io["P_FA"]
    fp_ = [(simp.get(PFA[k][0]) * Rinv % cfg.p, simp.get(PFA[k][1]) * Rinv % cfg.p) for k in range(6)]
    assert c.final_exp(fp_) == want
    assert simp.maxv < cfg.R
    simf = GM.Sim(gf)
    for k in range(6):
        for cc in range(2):
            simf.set(io["F_FA"][k][cc], sim.get(FA[k][cc]))
    simf.run("FINALEXP")
    OUT = io["OUT"]
    assert [(simf.get(OUT[k][0]) % cfg.p, simf.get(OUT[k][1]) % cfg.p) for k in range(6)] == want
    assert all(simf.get(OUT[k][cc]) < 2 * cfg.p for k in range(6) for cc in range(2))
    # Fp12 product programs
    a = [(rng.randrange(c.p), rng.randrange(c.p)) for _ in range(6)]
    b = [(rng.randrange(c.p), rng.randrange(c.p)) for _ in range(6)]
    sim2 = GM.Sim(gm)
    for k in range(6):
        for cc in range(2):
            sim2.set(FA[k][cc], a[k][cc] * cfg.R % cfg.p)
            sim2.set(io["GB"][k][cc], b[k][cc] * cfg.R % cfg.p)
    sim2.run("MUL_AB")
    FB = io["FB"]
    assert [(sim2.get(FB[k][0]) * Rinv % cfg.p, sim2.get(FB[k][1]) * Rinv % cfg.p) for k in range(6)] == c.fp12_mul(a, b)


def test_committed_tables_are_current(tmp_path):
    out = tmp_path / "machine_tables.cuh"
    GM.emit_tables(str(out))
    assert out.read_text() == open(os.path.join(ROOT, "bgls_b200", "csrc", "machine_tables.cuh")).read()
