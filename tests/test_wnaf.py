"""tools/gen_machine.py: the signed width-3 window of the final exponentiation's cyclotomic exponentiations -- digit
recoding, the exponentiation it describes (checked in a plain multiplicative group), and the phase counts it buys."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import gen_machine as GM  # noqa: E402


def wnaf3(e):
    out = []
    while e:
        if e & 1:
            d = e % 8
            if d > 4:
                d -= 8
            e -= d
        else:
            d = 0
        out.append(d)
        e >>= 1
    return out


def test_wnaf3_recoding_and_exponentiation():
    p = (1 << 127) - 1
    g = 3
    for e in (GM.BN_U, GM.BLS_X, (GM.BLS_X + 1) ** 2 // 3, 1, 3, 5, 7, 2 ** 64 - 1, 0x5555555555555555):
        d = wnaf3(e)
        assert sum(x << i for i, x in enumerate(d)) == e
        assert all(x in (0, 1, -1, 3, -3) for x in d) and d[-1] > 0
        assert all(not (d[i] and (d[i + 1] or (i + 2 < len(d) and d[i + 2]))) for i in range(len(d) - 2)), "two non-zero digits within a window"
        tbl = {k: pow(g, k, p) for k in (1, 3, -1, -3)}
        cur = tbl[d[-1]]
        for x in reversed(d[:-1]):
            cur = cur * cur % p
            if x:
                cur = cur * tbl[x] % p
        assert cur == pow(g, e, p)


def test_window_saves_phases():
    for cfg, limit in ((GM.BN, 610), (GM.BLS, 790)):
        gens, io = GM.build_all(cfg)
        assert len(gens["F"].programs["FINALEXP"]) <= limit
