"""Program generator for the slot engine (bgls_b200/csrc/slotvm.cuh).

The slot engine runs the Miller loops of the pairing product behind `CurveSystem.PairingProduct`
(/root/reference/curves/curve.go:125-170; altbn128.go:130-145; bls12_381.go:228-240) as straight-line programs of
Fp2 operations over a per-pair file of Fp2 *slots* in shared memory.  A pair is owned by G lanes (G = 1: one thread
per pair; G = 2: two lanes per pair, each round runs one operation per lane and the lanes synchronise afterwards).
All values are canonical Montgomery residues, so every operation is an exact field operation and a program can be
checked on a CPU by evaluating it over big integers (`Emu`, used by tests/test_slotvm_gen.py against the oracle).

Operations (32-bit words: kind | dst << 5 | a << 14 | b << 23; slots >= CONST0 are block-shared constants):
    MUL d = a*b   SQR d = a^2   ADD d = a+b   SUB d = a-b   XI d = xi*a   HALF d = a/2   CONJ d = conj(a)
    NEG d = -a    COPY d = a    NOP
    SEL0 / SEL1 d = flag(j) ? 0 / 1 : a   (j = b field: sub-pair of the group; flag = the pair has a point at infinity)

K pairs may share one Miller accumulator (K = 2 with 8 lanes per group): every iteration squares f once and multiplies
the K lines into it -- the squaring is 12 of the 37 Fp2 multiplications of a doubling iteration, and a group of K
pairs needs far fewer slots per pair, i.e. more resident warps.  A pair with a point at infinity must contribute 1:
its line is replaced by 1 through SEL0 / SEL1 (per-pair flags in shared memory).

Formulas (homogeneous projective doubling / mixed addition with lines, Karatsuba towers, sparse line
multiplication) are those of pairing.cuh / field.cuh, i.e. of oracle/pairing_impl.h, so the raw Miller value of the
binary loop equals the oracle's bit for bit; the default altbn128 loop uses the NAF of 6u+2 (21 instead of 36
addition steps), whose Miller value differs from the binary one by factors that the final exponentiation kills.

    python tools/gen_slotvm.py        # writes bgls_b200/csrc/slotvm_tables.cuh
"""
from __future__ import annotations

import os
import sys

NOP, MUL, SQR, ADD, SUB, XI, HALF, CONJ, NEG, COPY, SEL0, SEL1 = range(12)
KIND_NAMES = ["NOP", "MUL", "SQR", "ADD", "SUB", "XI", "HALF", "CONJ", "NEG", "COPY", "SEL0", "SEL1"]
HEAVY = (MUL, SQR)
COST = {MUL: 14, SQR: 11, ADD: 1, SUB: 1, XI: 3, HALF: 1, CONJ: 1, NEG: 1, COPY: 1, SEL0: 1, SEL1: 1, NOP: 0}
CONST0 = 256

BN_P = 21888242871839275222246405745257275088696311157297823662689037894645226208583
BN_U = 4965661367192848881
BLS_P = 0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab
BLS_X = 0xD201000000010000


def naf(n):
    out = []
    while n:
        if n & 1:
            d = 2 - (n % 4)
            n -= d
        else:
            d = 0
        out.append(d)
        n >>= 1
    return out[::-1]


class Field2:
    def __init__(self, p, xi):
        self.p, self.xi = p, xi

    def add(self, a, b):
        return ((a[0] + b[0]) % self.p, (a[1] + b[1]) % self.p)

    def sub(self, a, b):
        return ((a[0] - b[0]) % self.p, (a[1] - b[1]) % self.p)

    def mul(self, a, b):
        p = self.p
        return ((a[0] * b[0] - a[1] * b[1]) % p, (a[0] * b[1] + a[1] * b[0]) % p)

    def inv(self, a):
        p = self.p
        n = pow(a[0] * a[0] + a[1] * a[1], -1, p)
        return (a[0] * n % p, -a[1] * n % p)

    def pow(self, a, e):
        r = (1, 0)
        while e:
            if e & 1:
                r = self.mul(r, a)
            a = self.mul(a, a)
            e >>= 1
        return r


class Cfg:
    def __init__(self, name, cname, p, n, xi, twist, b, is_bn):
        self.name, self.cname, self.p, self.N, self.xi, self.twist, self.b, self.is_bn = name, cname, p, n, xi, twist, b, is_bn
        self.F = Field2(p, xi)
        self.R = 1 << (32 * n)
        F = self.F
        b2 = F.mul((b, 0), F.inv(xi)) if twist == "D" else F.mul((b, 0), xi)
        self.b2x3 = F.mul(b2, (3, 0))
        if is_bn:
            self.g12 = F.pow(xi, (p - 1) // 3)      # gamma_{1,2}
            self.g13 = F.pow(xi, (p - 1) // 2)      # gamma_{1,3}
            s = 6 * BN_U + 2
            self.loop_bin = [int(c) for c in bin(s)[3:]]
            self.loop_naf = naf(s)[1:]
        else:
            self.loop_bin = [int(c) for c in bin(BLS_X)[3:]]
            self.loop_naf = self.loop_bin


BN = Cfg("altbn128", "BN254", BN_P, 8, (9, 1), "D", 3, True)
BLS = Cfg("bls12-381", "BLS381", BLS_P, 12, (1, 1), "M", 4, False)


class Val:
    __slots__ = ("id", "kind", "a", "b", "state", "users", "prio", "round", "lane", "slot", "last", "imm")

    def __init__(self, id, kind, a=None, b=None, state=None, imm=0):
        self.id, self.kind, self.a, self.b, self.state, self.imm = id, kind, a, b, state, imm
        self.users = []
        self.slot = None


class Builder:
    """Straight-line program over SSA values; inputs / outputs are named state slots or constants."""

    def __init__(self, cfg: Cfg):
        self.cfg = cfg
        self.vals = []
        self.inputs = {}    # state name -> Val
        self.outputs = {}   # state name -> Val
        self.consts = {}    # const name -> Val

    def _new(self, kind, a=None, b=None, state=None, imm=0):
        v = Val(len(self.vals), kind, a, b, state, imm)
        self.vals.append(v)
        for x in (a, b):
            if x is not None:
                x.users.append(v)
        return v

    def inp(self, name):
        if name not in self.inputs:
            self.inputs[name] = self._new("in", state=name)
        return self.inputs[name]

    def const(self, name):
        if name not in self.consts:
            self.consts[name] = self._new("const", state=name)
        return self.consts[name]

    def out(self, name, v):
        self.outputs[name] = v

    # ---- Fp2 operations
    def mul(self, a, b): return self._new(MUL, a, b)
    def sqr(self, a): return self._new(SQR, a)
    def add(self, a, b): return self._new(ADD, a, b)
    def sub(self, a, b): return self._new(SUB, a, b)
    def xi(self, a): return self._new(XI, a)
    def half(self, a): return self._new(HALF, a)
    def conj(self, a): return self._new(CONJ, a)
    def neg(self, a): return self._new(NEG, a)
    def copy(self, a): return self._new(COPY, a)
    def sel0(self, a, j): return self._new(SEL0, a, imm=j)
    def sel1(self, a, j): return self._new(SEL1, a, imm=j)
    def dbl(self, a): return self.add(a, a)
    def triple(self, a): return self.add(self.add(a, a), a)

    # ---- Fp6 = Fp2[v]/(v^3 - xi): triples
    def f6_add(self, a, b): return tuple(self.add(x, y) for x, y in zip(a, b))
    def f6_sub(self, a, b): return tuple(self.sub(x, y) for x, y in zip(a, b))
    def f6_mul_v(self, a): return (self.xi(a[2]), a[0], a[1])

    def f6_mul(self, a, b):
        v0, v1, v2 = self.mul(a[0], b[0]), self.mul(a[1], b[1]), self.mul(a[2], b[2])
        u = self.mul(self.add(a[1], a[2]), self.add(b[1], b[2]))
        r0 = self.add(v0, self.xi(self.sub(self.sub(u, v1), v2)))
        u = self.mul(self.add(a[0], a[1]), self.add(b[0], b[1]))
        r1 = self.add(self.sub(self.sub(u, v0), v1), self.xi(v2))
        u = self.mul(self.add(a[0], a[2]), self.add(b[0], b[2]))
        r2 = self.add(self.sub(self.sub(u, v0), v2), v1)
        return (r0, r1, r2)

    def f6_mul_by_01(self, a, b0, b1):
        v0, v1 = self.mul(a[0], b0), self.mul(a[1], b1)
        r0 = self.add(v0, self.xi(self.mul(a[2], b1)))
        u = self.mul(self.add(a[0], a[1]), self.add(b0, b1))
        r1 = self.sub(self.sub(u, v0), v1)
        r2 = self.add(self.mul(a[2], b0), v1)
        return (r0, r1, r2)

    def f6_mul_by_0(self, a, b0): return tuple(self.mul(x, b0) for x in a)

    def f6_mul_by_1(self, a, b1):
        return (self.xi(self.mul(a[2], b1)), self.mul(a[0], b1), self.mul(a[1], b1))

    # ---- Fp12 = Fp6[w]/(w^2 - v): (c0, c1)
    def f12_sqr(self, f):
        a, b = f
        t = self.f6_mul(a, b)
        s = self.f6_mul(self.f6_add(a, b), self.f6_add(self.f6_mul_v(b), a))
        c0 = self.f6_sub(self.f6_sub(s, t), self.f6_mul_v(t))
        c1 = self.f6_add(t, t)
        return (c0, c1)

    def f12_mul(self, f, g):
        t0, t1 = self.f6_mul(f[0], g[0]), self.f6_mul(f[1], g[1])
        u = self.f6_mul(self.f6_add(f[0], f[1]), self.f6_add(g[0], g[1]))
        c1 = self.f6_sub(self.f6_sub(u, t0), t1)
        c0 = self.f6_add(t0, self.f6_mul_v(t1))
        return (c0, c1)

    def f12_mul_line(self, f, ly, lx, lc):
        """altbn128 (D): line = ly + lx w + lc w^3 -> ((ly,0,0),(lx,lc,0)); bls12-381 (M): ((lc,lx,0),(0,ly,0))."""
        if self.cfg.is_bn:
            t0 = self.f6_mul_by_0(f[0], ly)
            t1 = self.f6_mul_by_01(f[1], lx, lc)
            u = self.f6_mul_by_01(self.f6_add(f[0], f[1]), self.add(ly, lx), lc)
        else:
            t0 = self.f6_mul_by_01(f[0], lc, lx)
            t1 = self.f6_mul_by_1(f[1], ly)
            u = self.f6_mul_by_01(self.f6_add(f[0], f[1]), lc, self.add(lx, ly))
        c1 = self.f6_sub(self.f6_sub(u, t0), t1)
        c0 = self.f6_add(t0, self.f6_mul_v(t1))
        return (c0, c1)

    def line_as_f12(self, ly, lx, lc):
        z = self.const("ZERO")
        if self.cfg.is_bn:
            return ((ly, z, z), (lx, lc, z))
        return ((lc, lx, z), (z, ly, z))

    # ---- Miller steps (pairing.cuh: dbl_step / add_step)
    def dbl_step(self, T, px, py):
        X, Y, Z = T
        A = self.half(self.mul(X, Y))
        Bv = self.sqr(Y)
        Cv = self.sqr(Z)
        if self.cfg.is_bn:
            E = self.mul(Cv, self.const("B2X3"))
        else:   # 3 b' = 12 (1 + i): xi * C, then * 12
            t = self.xi(Cv)
            t3 = self.triple(t)
            E = self.dbl(self.dbl(t3))
        Fv = self.triple(E)
        Gv = self.half(self.add(Bv, Fv))
        H = self.dbl(self.mul(Y, Z))           # (Y+Z)^2 - B - C
        X2 = self.sqr(X)
        ly = self.mul(H, py)
        lx = self.neg(self.mul(self.triple(X2), px))
        lc = self.sub(Bv, E)
        X3 = self.mul(A, self.sub(Bv, Fv))
        Y3 = self.sub(self.sqr(Gv), self.triple(self.sqr(E)))
        Z3 = self.mul(Bv, H)
        return (X3, Y3, Z3), (ly, lx, lc)

    def add_step(self, T, Q, px, py):
        X, Y, Z = T
        qx, qy = Q
        th = self.sub(Y, self.mul(qy, Z))
        la = self.sub(X, self.mul(qx, Z))
        ly = self.mul(la, py)
        lx = self.neg(self.mul(th, px))
        lc = self.sub(self.mul(th, qx), self.mul(la, qy))
        Cv = self.sqr(th)
        D = self.sqr(la)
        E = self.mul(la, D)
        Fv = self.mul(Z, Cv)
        Gv = self.mul(X, D)
        H = self.sub(self.sub(self.add(E, Fv), Gv), Gv)
        X3 = self.mul(la, H)
        Y3 = self.sub(self.mul(th, self.sub(Gv, H)), self.mul(E, Y))
        Z3 = self.mul(Z, E)
        return (X3, Y3, Z3), (ly, lx, lc)


FSLOTS = (("F00", "F01", "F02"), ("F10", "F11", "F12"))
GSLOTS = (("G00", "G01", "G02"), ("G10", "G11", "G12"))
PAIR_STATE = ("TX", "TY", "TZ", "QX", "QY", "PX", "PY")
# F: Miller accumulator ((F00,F01,F02),(F10,F11,F12)) shared by the K pairs of a group; per pair j: T (projective point),
# Q (affine G2 input), PX = (xP, 0), PY = (yP, 0); G: second Fp12 operand of the product tree (aliases temporaries)


def state_names(k):
    return [n for h in FSLOTS for n in h] + ["%s%d" % (n, j) for j in range(k) for n in PAIR_STATE] + [n for h in GSLOTS for n in h]


def get_f(b, names=FSLOTS): return tuple(tuple(b.inp(n) for n in h) for h in names)
def get_t(b, j): return (b.inp("TX%d" % j), b.inp("TY%d" % j), b.inp("TZ%d" % j))


def put_f(b, f, names=FSLOTS):
    for h, hn in zip(f, names):
        for v, n in zip(h, hn):
            b.out(n, v)


def put_t(b, T, j):
    for v, n in zip(T, ("TX%d" % j, "TY%d" % j, "TZ%d" % j)):
        b.out(n, v)


def guard_line(b, line, j, k):
    """K > 1: the line of a pair with a point at infinity becomes 1 (the pair contributes 1 to the shared accumulator)."""
    if k == 1:
        return line   # a lone pair is fixed up after the loop (f <- 1)
    ly, lx, lc = line
    # 1 = ly + 0 w + 0 w^3 (D-type) / lc + 0 w^2 + 0 w^3 (M-type): the constant term is ly on altbn128, lc on bls12-381
    if b.cfg.is_bn:
        return (b.sel1(ly, j), b.sel0(lx, j), b.sel0(lc, j))
    return (b.sel0(ly, j), b.sel0(lx, j), b.sel1(lc, j))


def prog_dbl(cfg, k=1, first=False):
    """T_j <- 2 T_j, f <- f^2 * prod_j line_j (first: f == 1, so f <- prod_j line_j)."""
    b = Builder(cfg)
    lines = []
    for j in range(k):
        T, line = b.dbl_step(get_t(b, j), b.inp("PX%d" % j), b.inp("PY%d" % j))
        put_t(b, T, j)
        lines.append(guard_line(b, line, j, k))
    if first:
        f = b.line_as_f12(*lines[0])
        if k == 1:
            f = tuple(tuple(b.copy(x) for x in h) for h in f)
        rest = lines[1:]
    else:
        f = b.f12_sqr(get_f(b))
        rest = lines
    for line in rest:
        f = b.f12_mul_line(f, *line)
    if first and k > 1:   # outputs must be operation results (constants cannot be state)
        f = tuple(tuple(x if x.kind not in ("in", "const") else b.copy(x) for x in h) for h in f)
    put_f(b, f)
    return b


def prog_add(cfg, k=1, neg=False):
    b = Builder(cfg)
    f = get_f(b)
    for j in range(k):
        qy = b.inp("QY%d" % j)
        Q = (b.inp("QX%d" % j), b.neg(qy) if neg else qy)
        T, line = b.add_step(get_t(b, j), Q, b.inp("PX%d" % j), b.inp("PY%d" % j))
        put_t(b, T, j)
        f = b.f12_mul_line(f, *guard_line(b, line, j, k))
    put_f(b, f)
    return b


def prog_bn_frob(cfg, k, which):
    """which = 1: (QX, QY) <- Q1 = (conj(x) g2, conj(y) g3); which = 2: (QX, QY) <- -Q2 with Q2 = frobenius(Q1), i.e.
    (conj(x1) g2, -conj(y1) g3) -- pairing.cuh: miller_loop tail."""
    b = Builder(cfg)
    for j in range(k):
        x = b.mul(b.conj(b.inp("QX%d" % j)), b.const("G12"))
        y = b.mul(b.conj(b.inp("QY%d" % j)), b.const("G13"))
        if which == 2:
            y = b.neg(y)
        b.out("QX%d" % j, x)
        b.out("QY%d" % j, y)
    return b


def prog_conj(cfg):
    b = Builder(cfg)
    for n in FSLOTS[1]:
        b.out(n, b.neg(b.inp(n)))
    return b


def prog_mul12(cfg):
    """f <- f * g (product tree)."""
    b = Builder(cfg)
    put_f(b, b.f12_mul(get_f(b), get_f(b, GSLOTS)))
    return b


# ------------------------------------------------------------------------------------------- scheduling
class Scheduled:
    def __init__(self, cfg, g, rounds, nslots):
        self.cfg, self.g, self.rounds, self.nslots = cfg, g, rounds, nslots   # rounds: list of g-tuples of (kind, d, a, b)

    def words(self):
        out = []
        for r in self.rounds:
            for k, d, a, b in r:
                out.append(k | (d << 5) | (a << 14) | (b << 23))
        return out

    def cost(self):
        """Estimated warp time of the program in units of one light operation."""
        c = 0
        for r in self.rounds:
            kinds = {op[0] for op in r if op[0] != NOP}
            if kinds & set(HEAVY):
                c += COST[MUL] if MUL in kinds else COST[SQR]
            else:
                c += sum(COST[k] for k in kinds)
        return c

    def stats(self):
        heavy = sum(1 for r in self.rounds if any(op[0] in HEAVY for op in r))
        hops = sum(1 for r in self.rounds for op in r if op[0] in HEAVY)
        lops = sum(1 for r in self.rounds for op in r if op[0] not in HEAVY and op[0] != NOP)
        mixed = sum(1 for r in self.rounds if len({op[0] for op in r if op[0] != NOP}) > 1)
        return dict(rounds=len(self.rounds), heavy_rounds=heavy, heavy_ops=hops, light_ops=lops, mixed_rounds=mixed, slots=self.nslots)


def schedule(b: Builder, g: int, state_slots: dict, const_slots: dict, ntemp_base: int, alias_free=(), rng=None, order_w=None):
    """List scheduling into rounds of g lanes + slot allocation.  Returns Scheduled.
    state_slots: name -> slot; temporaries are allocated from ntemp_base upwards; the state slots named in
    `alias_free` hold no live value in this program and are handed to the allocator as extra temporaries."""
    ops = [v for v in b.vals if v.kind not in ("in", "const")]
    live_out = set(id(v) for v in b.outputs.values())
    # drop dead code
    needed = set()
    stack = list(b.outputs.values())
    while stack:
        v = stack.pop()
        if id(v) in needed:
            continue
        needed.add(id(v))
        for x in (v.a, v.b):
            if x is not None:
                stack.append(x)
    ops = [v for v in ops if id(v) in needed]
    for v in b.vals:
        v.users = [u for u in v.users if id(u) in needed]
    # priorities: longest path to a sink
    for v in reversed(ops):
        v.prio = COST[v.kind] + max((u.prio for u in v.users), default=0)
    if order_w is not None:   # lean towards program order (depth first: fewer live values)
        for v in ops:
            v.prio = v.prio * order_w - v.id
    if rng is not None:   # randomised tie-breaking / mild reordering: the caller keeps the schedule with the fewest slots
        for v in ops:
            v.prio = v.prio * (1.0 + 0.35 * rng.random())
    done = set(id(v) for v in b.vals if v.kind in ("in", "const"))
    remaining = list(ops)
    rounds_v = []
    while remaining:
        ready = [v for v in remaining if all(x is None or id(x) in done for x in (v.a, v.b))]
        ready.sort(key=lambda v: (-v.prio, v.id))
        first = ready[0]
        heavy = first.kind in HEAVY
        pool = [v for v in ready[1:] if (v.kind in HEAVY) == heavy]
        pool.sort(key=lambda v: (v.kind != first.kind, -v.prio, v.id))   # same kind first: the lanes of a warp then run one code path
        pick = [first] + pool[:g - 1]
        for v in pick:
            remaining.remove(v)
        rounds_v.append(pick)
        for v in pick:
            done.add(id(v))
    # ---- slot allocation (round granularity)
    for ri, r in enumerate(rounds_v):
        for li, v in enumerate(r):
            v.round, v.lane = ri, li
    nr = len(rounds_v)
    last_use = {}
    for v in b.vals:
        lu = -1
        for u in v.users:
            lu = max(lu, u.round)
        if id(v) in live_out:
            lu = nr          # live to the end
        last_use[id(v)] = lu
    out_target = {}
    for name, v in b.outputs.items():
        out_target.setdefault(id(v), []).append(name)
    slot_of = {}
    for name, v in b.inputs.items():
        slot_of[id(v)] = state_slots[name]
    for name, v in b.consts.items():
        slot_of[id(v)] = const_slots[name]
    # occupancy: slot -> round after which it is free (exclusive); state inputs occupy their slot until their last use
    busy_until = {}
    for name, v in b.inputs.items():
        busy_until[state_slots[name]] = last_use[id(v)]
    # state slots that are outputs but not inputs are free from the start; other state slots are not touched
    free_pool = []   # temporaries
    next_temp = [ntemp_base]
    extra = [state_slots[n] for n in alias_free]
    for s in extra:
        busy_until.setdefault(s, -1)

    def slot_free_at(s, r):   # may a value defined in round r live in slot s?
        return busy_until.get(s, -1) < r

    fixups = []
    state_set = set(state_slots.values()) - set(extra)
    for ri, r in enumerate(rounds_v):
        for v in r:
            tgt = out_target.get(id(v))
            chosen = None
            if tgt:
                s = state_slots[tgt[0]]
                if slot_free_at(s, ri) or _dies_here(b, v, s, slot_of, last_use, ri, r):
                    chosen = s
            if chosen is None:
                for s in extra + free_pool:
                    if slot_free_at(s, ri):
                        chosen = s
                        break
            if chosen is None:
                # in place over an operand that dies here (temporaries only: state slots are kept for their outputs)
                for x in (v.a, v.b):
                    if x is None or x.kind == "const":
                        continue
                    s = slot_of[id(x)]
                    if s not in state_set and _dies_here(b, v, s, slot_of, last_use, ri, r):
                        chosen = s
                        break
            if chosen is None:
                chosen = next_temp[0]
                next_temp[0] += 1
                free_pool.append(chosen)
            slot_of[id(v)] = chosen
            busy_until[chosen] = max(last_use[id(v)], ri)
            if tgt:
                for name in tgt:
                    if state_slots[name] != chosen:
                        fixups.append((name, v))
    # emit
    def enc(v, as_mul=False):
        k = v.kind
        a = slot_of[id(v.a)] if v.a is not None else 0
        bb = slot_of[id(v.b)] if v.b is not None else v.imm
        if as_mul:
            k, bb = MUL, a
        return (k, slot_of[id(v)], a, bb)

    rounds = []
    for r in rounds_v:
        kinds = {v.kind for v in r}
        conv = MUL in kinds and SQR in kinds
        ops_r = [enc(v, as_mul=(conv and v.kind == SQR)) for v in r]
        while len(ops_r) < g:
            ops_r.append((NOP, 0, 0, 0))
        rounds.append(tuple(ops_r))
    # copies into the state slots that could not be written in place (after everything else; sources stay live to the end)
    pend = [(state_slots[name], slot_of[id(v)]) for name, v in fixups]
    srcs = {s for _, s in pend}
    for d, _ in pend:
        assert d not in srcs, "copy cycle in state write-back"
    for i in range(0, len(pend), g):
        ops_r = [(COPY, d, s, 0) for d, s in pend[i:i + g]]
        while len(ops_r) < g:
            ops_r.append((NOP, 0, 0, 0))
        rounds.append(tuple(ops_r))
    check_hazards(rounds, g)
    return Scheduled(b.cfg, g, rounds, next_temp[0])


def _dies_here(b, v, s, slot_of, last_use, ri, rnd):
    """slot s holds an operand of v whose last use is v itself, and no other op of this round touches it."""
    holders = [x for x in (v.a, v.b) if x is not None and x.kind != "const" and slot_of.get(id(x)) == s]
    if not holders:
        return False
    for x in holders:
        if last_use[id(x)] != ri:
            return False
        if any(u is not v and u.round == ri for u in x.users):
            return False
    for o in rnd:
        if o is v:
            continue
        for x in (o.a, o.b):
            if x is not None and slot_of.get(id(x)) == s:
                return False
    return True


def check_hazards(rounds, g):
    for r in rounds:
        for i, (k, d, a, b) in enumerate(r):
            if k == NOP:
                continue
            for j, (k2, d2, a2, b2) in enumerate(r):
                if i == j or k2 == NOP:
                    continue
                assert d != d2, "two lanes write one slot"
                reads2 = {a2} | ({b2} if k2 in (MUL, ADD, SUB) else set())
                assert d not in reads2, "lane writes a slot another lane reads in the same round"


# ------------------------------------------------------------------------------------------- whole engine description
class Engine:
    """All programs of one curve for one group shape (g lanes per group of k pairs), with a common slot map."""

    def __init__(self, cfg: Cfg, g: int, k: int = 1, tries=1, slot_weight=0.0):
        self.cfg, self.g, self.k = cfg, g, k
        names = state_names(k)
        nstate = 6 + 7 * k
        self.state = {n: i for i, n in enumerate(names)}   # F, per-pair state, then G right after (fixed: the tree's copy-in code uses them)
        pair_state = names[6:nstate]
        gnames = names[nstate:]
        self.const_names = ["ZERO", "ONE"] + (["B2X3", "G12", "G13"] if cfg.is_bn else [])
        self.consts = {n: CONST0 + i for i, n in enumerate(self.const_names)}
        self.progs = {}
        ntemp_g = nstate + 6
        specs = [("DBL", lambda: prog_dbl(cfg, k), gnames), ("DBL1", lambda: prog_dbl(cfg, k, first=True), gnames),
                 ("ADD", lambda: prog_add(cfg, k), gnames), ("CONJ", lambda: prog_conj(cfg), ()), ("MUL12", lambda: prog_mul12(cfg), pair_state)]
        if cfg.is_bn:
            specs += [("SUBQ", lambda: prog_add(cfg, k, neg=True), gnames),
                      ("FROB1", lambda: prog_bn_frob(cfg, k, 1), ()), ("FROB2", lambda: prog_bn_frob(cfg, k, 2), ())]
        self.nslots = ntemp_g
        import random
        for name, mk, alias in specs:
            best = None
            for t in range(tries):
                rng = random.Random(1000 * t + 7) if t >= 3 else None
                order_w = (None, 0.0, 0.5)[t % 3]
                s = schedule(mk(), g, self.state, self.consts, ntemp_g, alias_free=alias, rng=rng, order_w=order_w)
                key = (s.cost() + slot_weight * s.nslots, s.nslots)
                if best is None or key < best[0]:
                    best = (key, s)
            self.progs[name] = best[1]
            self.nslots = max(self.nslots, best[1].nslots)

    def const_values(self):
        cfg = self.cfg
        vals = {"ZERO": (0, 0), "ONE": (1, 0)}
        if cfg.is_bn:
            vals.update({"B2X3": cfg.b2x3, "G12": cfg.g12, "G13": cfg.g13})
        return [vals[n] for n in self.const_names]

    def sequence(self, use_naf=True):
        """Program names of one Miller loop, in order (f = 1, T = Q on entry)."""
        cfg = self.cfg
        digits = cfg.loop_naf if (use_naf and cfg.is_bn) else cfg.loop_bin
        seq = []
        for i, d in enumerate(digits):
            seq.append("DBL1" if i == 0 else "DBL")
            if d == 1:
                seq.append("ADD")
            elif d == -1:
                seq.append("SUBQ")
        if cfg.is_bn:
            seq += ["FROB1", "ADD", "FROB2", "ADD"]
        else:
            seq.append("CONJ")
        return seq


class Emu:
    """Exact evaluation of scheduled programs over big integers (plain residues, not Montgomery)."""

    def __init__(self, eng: Engine):
        self.eng = eng
        self.F = eng.cfg.F
        self.slots = {}
        self.flags = 0       # bit j: sub-pair j of the group has a point at infinity
        for s, v in zip(range(CONST0, CONST0 + len(eng.const_names)), eng.const_values()):
            self.slots[s] = v

    def set(self, name, v):
        self.slots[self.eng.state[name]] = (v[0] % self.F.p, v[1] % self.F.p)

    def get(self, name):
        return self.slots[self.eng.state[name]]

    def run(self, prog):
        F, p = self.F, self.F.p
        inv2 = pow(2, -1, p)
        for r in self.eng.progs[prog].rounds:
            res = []
            for k, d, a, b in r:   # all lanes read before any lane writes (hazard-checked by the generator)
                if k == NOP:
                    continue
                x = self.slots[a]
                if k == MUL:
                    v = F.mul(x, self.slots[b])
                elif k == SQR:
                    v = F.mul(x, x)
                elif k == ADD:
                    v = F.add(x, self.slots[b])
                elif k == SUB:
                    v = F.sub(x, self.slots[b])
                elif k == XI:
                    v = F.mul(x, F.xi)
                elif k == HALF:
                    v = (x[0] * inv2 % p, x[1] * inv2 % p)
                elif k == CONJ:
                    v = (x[0], -x[1] % p)
                elif k == NEG:
                    v = (-x[0] % p, -x[1] % p)
                elif k == COPY:
                    v = x
                elif k in (SEL0, SEL1):
                    v = ((1, 0) if k == SEL1 else (0, 0)) if (self.flags >> b) & 1 else x
                else:
                    raise ValueError(k)
                res.append((d, v))
            for d, v in res:
                self.slots[d] = v

    def miller(self, Ps, Qs, use_naf=True):
        """Ps[j] = (x, y) ints or None (infinity), Qs[j] = ((x0, x1), (y0, y1)) or None, j < K; returns the shared
        accumulator f = prod_j f_j as [w^0..w^5] Fp2 coefficients (K = 1: the caller replaces f by 1 for an infinity pair)."""
        if self.eng.k == 1 and not isinstance(Ps, list):
            Ps, Qs = [Ps], [Qs]
        for n in FSLOTS[0] + FSLOTS[1]:
            self.set(n, (0, 0))
        self.set("F00", (1, 0))
        self.flags = 0
        for j, (P, Q) in enumerate(zip(Ps, Qs)):
            if P is None or Q is None:
                self.flags |= 1 << j
                P, Q = (0, 0), ((0, 0), (0, 0))
            self.set("TX%d" % j, Q[0]); self.set("TY%d" % j, Q[1]); self.set("TZ%d" % j, (1, 0))
            self.set("QX%d" % j, Q[0]); self.set("QY%d" % j, Q[1])
            self.set("PX%d" % j, (P[0], 0)); self.set("PY%d" % j, (P[1], 0))
        for name in self.eng.sequence(use_naf):
            self.run(name)
        c0 = [self.get(n) for n in FSLOTS[0]]
        c1 = [self.get(n) for n in FSLOTS[1]]
        return [c0[0], c1[0], c0[1], c1[1], c0[2], c1[2]]


# ------------------------------------------------------------------------------------------- emission
def words32(v, n):
    return [(v >> (32 * i)) & 0xFFFFFFFF for i in range(n)]


def emit(path):
    out = ["// GENERATED by tools/gen_slotvm.py -- do not edit.", "// Programs and constants of the slot engine (slotvm.cuh).",
           "#pragma once", "#include <cstdint>", "namespace bgls { namespace svt {"]
    for cfg in (BN, BLS):
        for g, kk in SHAPES:
            eng = engine(cfg, g, kk)
            tag = "%s_G%d" % (cfg.cname, g) + ("K%d" % kk if kk > 1 else "")
            seq = eng.sequence(True)
            names = list(eng.progs)
            out.append("struct %s {" % tag)
            out.append("    static constexpr int G = %d, K = %d, NSLOT = %d, NCONST = %d, NPROG = %d, SEQ_LEN = %d;" %
                       (g, kk, eng.nslots, len(eng.const_names), len(names), len(seq)))
            for n, s in eng.state.items():
                out.append("    static constexpr int S_%s = %d;" % (n, s))
            for i, n in enumerate(names):
                out.append("    static constexpr int P_%s = %d;" % (n, i))
            offs, words = [], []
            for n in names:
                offs.append(len(words))
                words += eng.progs[n].words()
            offs.append(len(words))
            out.append("    static constexpr int NWORDS = %d;" % len(words))
            out.append("    static const uint32_t* code() { static const uint32_t v[%d] = {%s}; return v; }" %
                       (len(words), ", ".join("0x%08xu" % w for w in words)))
            out.append("    static const uint32_t* offsets() { static const uint32_t v[%d] = {%s}; return v; }" %
                       (len(offs), ", ".join(str(o) for o in offs)))
            out.append("    static const uint8_t* sequence() { static const uint8_t v[%d] = {%s}; return v; }" %
                       (len(seq), ", ".join(str(names.index(s)) for s in seq)))
            cw = []
            for v in eng.const_values():
                for c in v:
                    cw += words32(c * cfg.R % cfg.p, cfg.N)
            out.append("    static const uint32_t* consts() { static const uint32_t v[%d] = {%s}; return v; }" %
                       (len(cw), ", ".join("0x%08xu" % w for w in cw)))
            ml = 10 if cfg.is_bn else 14   # limbs of the machine's 28-bit form (tools/gen_machine.py)
            out.append("    static constexpr int MACH_L = %d;" % ml)
            out.append("    static const uint32_t* mach_r() { static const uint32_t v[%d] = {%s}; return v; }" %
                       (cfg.N, ", ".join("0x%08xu" % w for w in words32(pow(2, 28 * ml, cfg.p), cfg.N))))
            out.append("};")
            st = {n: eng.progs[n].stats() for n in names}
            out.append("// %s: slots %d; %s" % (tag, eng.nslots, "; ".join("%s %d rounds (%d heavy ops, %d light, %d mixed)" % (
                n, s["rounds"], s["heavy_ops"], s["light_ops"], s["mixed_rounds"]) for n, s in st.items())))
    out.append("} }  // namespace bgls::svt")
    with open(path, "w") as f:
        f.write("\n".join(out) + "\n")


LANES = (1, 2, 4)
SHAPES = ((1, 1), (2, 1), (4, 1), (8, 1), (16, 1), (8, 2), (4, 2))   # (lanes per group, pairs per group)
_ENGINES = {}


def engine(cfg, g, k=1):
    """The released engine description of (curve, lanes per group, pairs per group): schedule search with a mild
    preference for fewer slots."""
    key = (cfg.name, g, k)
    if key not in _ENGINES:
        _ENGINES[key] = Engine(cfg, g, k, tries=120, slot_weight=4.0)
    return _ENGINES[key]


if __name__ == "__main__":
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    emit(os.path.join(root, "bgls_b200", "csrc", "slotvm_tables.cuh"))
    for cfg in (BN, BLS):
        for g, kk in SHAPES:
            eng = engine(cfg, g, kk)
            print(cfg.name, "G =", g, "K =", kk, "slots", eng.nslots, {n: s.stats() for n, s in eng.progs.items()}, file=sys.stderr)
