"""Schedule generator for the warp-cooperative "dot-product machine" (bgls_b200/csrc/machine.cuh).

One *group* of G = 16 lanes owns one pairing (or one Fp12 value).  All state lives in a per-group
file of Fp *slots* in shared memory; a program is a list of *phases*; in a phase every lane runs at
most one task of the same kind and the group synchronises afterwards:

  DOT  dst = MontRed( sum_t  slot[a_t] * slot[b_t] )        (lazy reduction: one reduction per output)
  LIN  dst = Normalize( sum_t  c_t * slot[s_t]   or   |c_t| * (KP_K - slot[s_t]) )   (small integer c_t)

Fp elements are L limbs of W = 28 bits (altbn128: L = 10, bls12-381: L = 14) in Montgomery form with
R = 2^(W L); values are only bounded (never conditionally reduced): R/p >= 2^11 gives the headroom
(altbn128 needs the 10th limb: xi = 9+i inflates lazy bounds ~10x per product, 9 x 29 bits only leaves 2^7).

Fp12 = Fp2[w]/(w^6 - xi), Fp2 = Fp[i]/(i^2+1), coordinates (k, re/im), k = power of w.  An Fp12
product is 12 dot products of 12 terms after the second operand has been expanded into its "derived
copies" (-y, xi*(x+iy), -(xi*..).y) by LIN tasks.

This file only *generates and simulates* (exact big-integer semantics, used by tests/ to prove the
programs against the oracle on a CPU); it never runs in the product path.  Output:
bgls_b200/csrc/machine_tables.cuh.       python tools/gen_machine.py
"""
from __future__ import annotations

import os

G = 16  # lanes per group
MAXT = 12


class Cfg:
    def __init__(self, name, p, r, W, L, xi_a, twist, b, loop, loop_naf, u, extra=None):
        self.name, self.p, self.r, self.W, self.L = name, p, r, W, L
        self.xi_a, self.twist, self.b = xi_a, twist, b
        self.loop, self.loop_naf, self.u = loop, loop_naf, u
        self.R = 1 << (W * L)
        self.mask = (1 << W) - 1
        self.n0 = (-pow(p, -1, 1 << W)) % (1 << W)
        self.extra = extra or {}

    def mont(self, v):
        return v % self.p * self.R % self.p

    def limbs(self, v):
        assert 0 <= v < self.R
        return [(v >> (self.W * i)) & self.mask for i in range(self.L)]

    def kp_limbs(self, K):
        """K*p as a limb vector whose low limbs are all >= 2^W - 1 (limb-wise subtraction never borrows)."""
        v = self.limbs(K * self.p)
        W, L = self.W, self.L
        out = [v[0] + (1 << W)] + [v[i] + (1 << W) - 1 for i in range(1, L - 1)] + [v[L - 1] - 1]
        assert sum(x << (W * i) for i, x in enumerate(out)) == K * self.p and out[-1] >= 0
        return out


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


BN_U = 4965661367192848881
BN = Cfg("BN254", 21888242871839275222246405745257275088696311157297823662689037894645226208583,
         21888242871839275222246405745257275088548364400416034343698204186575808495617,
         28, 10, 9, "D", 3, 6 * BN_U + 2, True, BN_U)
BLS_X = 0xD201000000010000
BLS = Cfg("BLS381", 0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab,
          52435875175126190479447740508185965837690552500527637822603658699938581184513,
          28, 14, 1, "M", 4, BLS_X, False, BLS_X)

KP_SET = sorted({m << k for k in range(0, 15) for m in (2, 3)})


def f2mul(a, b, p):
    return ((a[0] * b[0] - a[1] * b[1]) % p, (a[0] * b[1] + a[1] * b[0]) % p)


def f2pow(a, e, p):
    r = (1, 0)
    while e:
        if e & 1:
            r = f2mul(r, a, p)
        a = f2mul(a, a, p)
        e >>= 1
    return r


def f2inv(a, p):
    n = pow(a[0] * a[0] + a[1] * a[1], -1, p)
    return (a[0] * n % p, -a[1] * n % p)


class Phase:
    def __init__(self, kind):
        self.kind = kind  # "DOT" | "LIN" | "INV" (single-lane modular inverse of a plain-domain value)
        self.tasks = []   # (dst, terms)  DOT terms: (a, b); LIN terms: (slot, coef, K)

    @property
    def T(self):
        return max((len(t[1]) for t in self.tasks), default=0)


class Gen:
    """Builds phases/programs for one curve and tracks static value bounds."""

    def __init__(self, cfg: Cfg, lanes=G, tm=MAXT, split=False, signed=False):
        self.cfg = cfg
        self.lanes, self.tm, self.split = lanes, tm, split   # lanes per group, max terms per DOT record, split long dots
        # signed slot files: a DOT term may carry a minus sign (the interpreter negates the limbs of its first
        # operand and accumulates in signed 64-bit columns); the phase header holds an offset K so that K p R
        # dominates the negative part, i.e. the reduced value (T + m p)/R + K p is never negative.
        self.signed = signed
        if signed:
            assert (tm + 1) * cfg.L * (1 << (2 * cfg.W)) < (1 << 63), "signed 64-bit columns would overflow"
        self.gslots = {}       # name -> index (per-group slot)
        self._nslots, self._pool, self._pools, self._pool_used = 0, None, {}, {}
        self.cslots = {}       # name -> index (block-shared constant slot)
        self.cvals = []        # constant limb vectors
        self.cints = []        # constant integer values (for the simulator)
        self.phases = []
        self.programs = {}     # name -> list of phase ids
        self.ub = {}           # slot ref -> static upper bound (integer)
        self.cur = None        # program being built
        self.cache = {}
        p = cfg.p
        self.ZERO = self.const("ZERO", 0, raw=True)
        self.ONE = self.const("ONE", 1)
        self.RAW1 = self.const("RAW1", 1, raw=True)
        self.R2 = self.const("R2", cfg.R * cfg.R % p, raw=True)
        self.KP = {}
        for K in KP_SET:
            if 2 * K * p >= cfg.R:
                break
            if signed:
                # signed files never load the borrow-free K p vectors (one offset constant per LIN task instead):
                # K stays a choice of the bound analysis only
                self.KP[K] = ("kp", K)
                continue
            idx = len(self.cvals)
            self.cslots["KP%d" % K] = idx
            self.cvals.append(cfg.kp_limbs(K))
            self.cints.append(K * p)
            self.KP[K] = ("c", idx)
            self.ub[("c", idx)] = None  # never a DOT operand

    # ---- slots
    def g(self, name):
        if name not in self.gslots:
            if self._pool is not None:
                # temporaries of code regions that are never live at the same time share physical slots:
                # the k-th new name of each region maps to the k-th slot of the pool
                pool, ns = self._pool
                k = self._pool_used.get((pool, ns), 0)
                self._pool_used[(pool, ns)] = k + 1
                slots = self._pools.setdefault(pool, [])
                if k == len(slots):
                    slots.append(self._nslots)
                    self._nslots += 1
                self.gslots[name] = slots[k]
            else:
                self.gslots[name] = self._nslots
                self._nslots += 1
        return ("g", self.gslots[name])

    def pool(self, pool, ns):
        gen = self

        class _P:
            def __enter__(self_):
                self_.prev = gen._pool
                gen._pool = (pool, ns)

            def __exit__(self_, *a):
                gen._pool = self_.prev
        return _P()

    def fp2(self, name):
        return (self.g(name + ".x"), self.g(name + ".y"))

    def const(self, name, value, raw=False):
        if name in self.cslots:
            return ("c", self.cslots[name])
        v = value % self.cfg.p if raw else self.cfg.mont(value)
        idx = len(self.cvals)
        self.cslots[name] = idx
        self.cvals.append(self.cfg.limbs(v))
        self.cints.append(v)
        self.ub[("c", idx)] = v
        return ("c", idx)

    def offset_const(self, k):
        """Constant slot holding the plain limbs of k * p (not reduced; never a DOT operand)."""
        if k == 0:
            return self.ZERO
        name = "KT%d" % k
        if name not in self.cslots:
            v = k * self.cfg.p
            assert v < self.cfg.R
            idx = len(self.cvals)
            self.cslots[name] = idx
            self.cvals.append(self.cfg.limbs(v))
            self.cints.append(v)
            self.ub[("c", idx)] = None
        return ("c", self.cslots[name])

    def const2(self, name, val2):
        return (self.const(name + ".x", val2[0]), self.const(name + ".y", val2[1]))

    def set_ub(self, slot, ub):
        self.ub[slot] = ub

    # ---- phases
    def begin(self, prog):
        self.cur = self.programs.setdefault(prog, [])

    def emit(self, phase: Phase, key=None):
        """Append the phase to the current program (re-using an identical earlier phase)."""
        cfg = self.cfg
        assert 0 < len(phase.tasks) <= self.lanes, len(phase.tasks)
        dsts = [t[0] for t in phase.tasks]
        assert len(set(dsts)) == len(dsts), "two tasks write one slot"
        reads = set()
        for dst, terms in phase.tasks:
            assert dst[0] == "g"
            for t in terms:
                reads.add(t[0])
                if phase.kind == "DOT":
                    reads.add(t[1])
        if phase.kind == "INV":
            assert len(phase.tasks) == 1 and len(phase.tasks[0][1]) == 1
        elif phase.kind == "DOT":
            assert not (reads & set(dsts)), "DOT phase reads a slot it writes"
            assert phase.T <= self.tm
        else:
            # a LIN task may overwrite its own inputs (it reads everything before writing) but not
            # another task's inputs
            for dst, terms in phase.tasks:
                others = set()
                for d2, t2 in phase.tasks:
                    if d2 != dst:
                        others |= {t[0] for t in t2}
                assert dst not in others, "LIN phase: cross-lane read/write hazard"
            assert phase.T <= 4
        # static bounds
        new_ub = {}
        for dst, terms in phase.tasks:
            if phase.kind == "INV":
                assert self.ub.get(terms[0][0]) is not None and self.ub[terms[0][0]] < 2 * cfg.p, "INV input must be < 2p"
                new_ub[dst] = cfg.p - 1
                continue
            if phase.kind == "DOT":
                s, neg = 0, 0
                for t in terms:
                    a, b = t[0], t[1]
                    assert self.ub.get(a) is not None and self.ub.get(b) is not None, ("unbounded DOT operand", a, b)
                    if len(t) > 2 and t[2] < 0:
                        assert self.signed, "negative DOT term in an unsigned slot file"
                        neg += self.ub[a] * self.ub[b]
                    else:
                        s += self.ub[a] * self.ub[b]
                if neg:
                    need = -(-(neg + neg // 2) // (cfg.p * cfg.R))   # 1.5x margin: phases are re-used
                    phase.K = max(getattr(phase, "K", 0), need)
                ub = s // cfg.R + cfg.p    # + K p, added below once the phase's K is known
            else:
                ub = 0
                for slot, coef, K in terms:
                    assert self.ub.get(slot) is not None, ("LIN operand without bound", slot)
                    if coef > 0:
                        ub += coef * self.ub[slot]
                    else:
                        assert K in KP_SET
                        kp = cfg.kp_limbs(K)
                        top = self.ub[slot] >> (cfg.W * (cfg.L - 1))
                        assert kp[-1] >= top, ("KP too small", K, slot)
                        ub += -coef * K * cfg.p
            assert ub < cfg.R, ("value bound exceeds R", ub.bit_length())
            new_ub[dst] = ub
        if phase.kind == "LIN" and self.signed:
            # signed files: the interpreter accumulates sum c_t * s_t with signed c_t and adds ONE constant per task,
            # (sum over the negative terms of |c_t| K_t) * p -- the same value as the term-wise K_t p - s_t form
            phase.kt = []
            for dst, terms in phase.tasks:
                kt = sum(-coef * K for _, coef, K in terms if coef < 0)
                phase.kt.append(self.offset_const(kt))
        if phase.kind == "DOT" and getattr(phase, "K", 0):
            assert phase.K < 256
            for dst in new_ub:
                new_ub[dst] += phase.K * cfg.p
                assert new_ub[dst] < cfg.R
        sig = (phase.kind, tuple((d, tuple(t)) for d, t in phase.tasks), getattr(phase, "nplain", None))
        if sig in self.cache:
            pid = self.cache[sig]   # bounds of every execution are re-verified by verify_program()
            if getattr(phase, "K", 0) > getattr(self.phases[pid], "K", 0):
                self.phases[pid].K = phase.K
        else:
            pid = len(self.phases)
            self.phases.append(phase)
            self.cache[sig] = pid
        self.ub.update(new_ub)
        self.cur.append(pid)
        return pid

    def verify_program(self, prog, init_ub):
        """Exact static bound check of one program in execution order: every operand is defined, every value
        stays below R, every KP constant dominates its subtrahend limb-wise, DOT operands are normalised."""
        cfg = self.cfg
        ub = {("c", i): (self.cints[i] if self.ub.get(("c", i)) is not None else None) for i in range(len(self.cvals))}
        kpc = {v: k for k, v in self.KP.items()}
        ub.update(init_ub)
        worst = 0
        for pid in self.programs[prog]:
            ph = self.phases[pid]
            new = {}
            for dst, terms in ph.tasks:
                if ph.kind == "INV":
                    assert ub.get(terms[0][0]) is not None and ub[terms[0][0]] < 2 * cfg.p, ("INV input must be < 2p", prog, pid)
                    new[dst] = cfg.p - 1
                    continue
                if ph.kind == "DOT":
                    acc, neg = 0, 0
                    for t in terms:
                        a, b = t[0], t[1]
                        assert ub.get(a) is not None and ub.get(b) is not None, ("undefined / non-normalised DOT operand", prog, pid)
                        if len(t) > 2 and t[2] < 0:
                            neg += ub[a] * ub[b]
                        else:
                            acc += ub[a] * ub[b]
                    K = getattr(ph, "K", 0)
                    assert neg <= K * cfg.p * cfg.R, ("offset K too small for the negative part", prog, pid, K)
                    v = acc // cfg.R + cfg.p + K * cfg.p
                else:
                    v = 0
                    for slot, coef, K in terms:
                        assert ub.get(slot) is not None, ("undefined LIN operand", prog, pid, slot)
                        if coef > 0:
                            v += coef * ub[slot]
                        else:
                            assert cfg.kp_limbs(K)[-1] >= (ub[slot] >> (cfg.W * (cfg.L - 1))), ("KP too small", prog, pid)
                            v += -coef * K * cfg.p
                assert v < cfg.R, ("bound exceeds R", prog, pid)
                worst = max(worst, v)
                new[dst] = v
            ub.update(new)
        return worst, ub

    def inv(self, dst, src):
        """dst = src^-1 in Montgomery form (src in Montgomery form, any bounded representative):
        out of Montgomery form (DOT by the plain 1), binary extended GCD on one lane, back (DOT by R^2)."""
        t0, t1 = self.g("INVOP.t0"), self.g("INVOP.t1")
        self.dot([(t0, [(src, self.RAW1)])])
        ph = Phase("INV")
        ph.tasks = [(t1, [(t0,)])]
        self.emit(ph)
        self.dot([(dst, [(t1, self.R2)])])

    def lin_rounds(self, tasks):
        """LIN tasks, 16 per phase (a later round must not read what an earlier round wrote)."""
        n = self.lanes
        for i in range(n, len(tasks), n):
            written = {t[0] for t in tasks[:i]}
            for d, terms in tasks[i:i + n]:
                assert not ({t[0] for t in terms} & (written - {d})), "LIN rounds: read after write across rounds"
        for i in range(0, len(tasks), n):
            ph = Phase("LIN")
            ph.tasks = tasks[i:i + n]
            self.emit(ph)

    def dot(self, tasks):
        """One DOT phase.  In split mode (32-lane groups) every dot of >= 4 terms is computed as two half dots on
        two lanes and summed by a following LIN phase: halves the dependent multiply chain of the phase."""
        ph = Phase("DOT")
        if not self.split:
            ph.tasks = tasks
            self.emit(ph)
            return
        comb = []
        for i, (dst, terms) in enumerate(tasks):
            if len(terms) >= 4:
                h = (len(terms) + 1) // 2
                a, b = self.g("SPL.a%d" % i), self.g("SPL.b%d" % i)
                ph.tasks += [(a, terms[:h]), (b, terms[h:])]
                comb.append((dst, a, b))
            else:
                ph.tasks.append((dst, terms))
        self.emit(ph)
        if comb:
            self.lin_rounds([self.lin(dst, (a, 1), (b, 1)) for dst, a, b in comb])

    def kdot(self, triples, plain):
        """DOT phase in Karatsuba-lane form.  `triples`: [(re_dst, im_dst, [(a, b), ...])] with a, b Fp2 slot pairs: the
        Fp2 sum of products sum_t a_t * b_t is computed by THREE lanes,  Q = sum a.y b.y,  P = sum a.x b.x,
        S = sum (a.x + a.y)(b.x + b.y)  (operand sums formed on the fly, at most tm/2 terms), which the interpreter
        combines by warp shuffle into re = P - Q and im = S - P - Q before the reduction.  `plain`: ordinary DOT
        tasks of at most tm/2 terms on the remaining lanes.  For bounds, hazards and the exact simulator the phase is
        the equivalent schoolbook DOT phase; the lane form only exists in the emitted tables."""
        assert self.signed and 3 * len(triples) + len(plain) <= self.lanes
        cfg = self.cfg
        half = self.tm // 2
        assert self.kdot_fits(), "K-lane columns would overflow"
        ph = Phase("DOT")
        ph.triples, ph.nplain = [], len(plain)
        for dst, terms in plain:
            assert len(terms) <= half
        ph.tasks = list(plain)
        for re, im, terms in triples:
            assert 0 < len(terms) <= half
            terms = [(t[0], t[1], t[2] if len(t) > 2 else 1) for t in terms]   # (a, b, sign of the Fp2 product)
            ph.triples.append((re, im, terms))
            ph.tasks += [(re, sum(([(a[0], b[0], sg), (a[1], b[1], -sg)] for a, b, sg in terms), [])),
                         (im, sum(([(a[0], b[1], sg), (a[1], b[0], sg)] for a, b, sg in terms), []))]
        self.emit(ph)

    def kdot_fits(self):
        """64-bit columns of a K-lane phase.  The three partial sums are accumulated modulo 2^64 (unsigned, exact), so
        S = sum (a.x+a.y)(b.x+b.y) may wrap; what has to fit the signed range is what the reduction sees after the
        combination: |P - Q| and |S - P - Q| = |sum a.x b.y + a.y b.x|, at most tm/2 terms of 2 L products of W-bit
        limbs, plus the L products the reduction adds per column and the offset K p."""
        cfg, half = self.cfg, self.tm // 2
        worst = half * 2 * cfg.L * (1 << (2 * cfg.W)) + cfg.L * (1 << (2 * cfg.W)) + (1 << 40)
        return self.signed and worst < (1 << 63)

    # ---- LIN helpers (return task tuples)
    def kfor(self, slot):
        top = (self.ub[slot] + self.ub[slot] // 4) >> (self.cfg.W * (self.cfg.L - 1))   # 1.25x margin: phases are re-used
        for K in sorted(self.KP):
            if self.cfg.kp_limbs(K)[-1] >= top:
                return K
        names = {v: k for k, v in self.gslots.items()}
        raise AssertionError("no KP constant large enough for %s (ub = %.1f p)" % (names.get(slot[1]), self.ub[slot] / self.cfg.p))

    def T(self, slot, coef):
        """LIN term."""
        return (slot, coef, self.kfor(slot) if coef < 0 else 0)

    def lin(self, dst, *pairs):
        return (dst, [self.T(s, c) for s, c in pairs])

    # xi * (x + i y) = (a x - y) + (x + a y) i
    def derived(self, v, d):
        """LIN tasks filling d = {ny, tx, ty, nty} from the Fp2 v = (x, y)."""
        a = self.cfg.xi_a
        x, y = v
        return [self.lin(d["ny"], (y, -1)), self.lin(d["tx"], (x, a), (y, -1)), self.lin(d["ty"], (x, 1), (y, a)),
                self.lin(d["nty"], (x, -1), (y, -a))]

    def der_slots(self, name):
        return {k: self.g(name + "." + k) for k in ("ny", "tx", "ty", "nty")}

    # Fp2 product u*v as two DOT tasks; nvy = slot holding -v.y
    def mul2(self, dst, u, v, nvy):
        return [(dst[0], [(u[0], v[0]), (u[1], nvy)]), (dst[1], [(u[0], v[1]), (u[1], v[0])])]

    def mul2_fp(self, dst, u, s):
        return [(dst[0], [(u[0], s)]), (dst[1], [(u[1], s)])]

    # ---- Fp12 blocks.  An Fp12 register is a list of 6 Fp2 (k = power of w).
    def f12(self, name):
        return [self.fp2("%s.%d" % (name, k)) for k in range(6)]

    def f12_derived(self, b, dname="DER"):
        ds = [self.der_slots("%s.%d" % (dname, j)) for j in range(6)]
        tasks = []
        for j in range(6):
            tasks += self.derived(b[j], ds[j])
        self.lin_rounds(tasks)
        return ds

    def f12_mul(self, d, a, b, bder=None):
        """d = a*b (d distinct from a, b)."""
        assert d is not a and d is not b
        ds = bder if bder is not None else self.f12_derived(b)
        tasks = []
        for k in range(6):
            re, im = [], []
            for i in range(6):
                j = (k - i) % 6
                wrapped = i > k
                bx, by = b[j]
                if not wrapped:
                    re += [(a[i][0], bx), (a[i][1], ds[j]["ny"])]
                    im += [(a[i][0], by), (a[i][1], bx)]
                else:
                    re += [(a[i][0], ds[j]["tx"]), (a[i][1], ds[j]["nty"])]
                    im += [(a[i][0], ds[j]["ty"]), (a[i][1], ds[j]["tx"])]
            tasks += [(d[k][0], re), (d[k][1], im)]
        self.dot(tasks)

    def f12_sqr(self, d, a):
        """d = a^2 (d distinct from a) by complex squaring over Fp6 = Fp2[v]/(v^3 - xi), v = w^2:
             a = c0 + c1 w,  c0 = (a0, a2, a4), c1 = (a1, a3, a5)
             t = c0 c1,  s = (c0 + c1)(c0 + v c1)   ->   a^2 = (s - t - v t) + (2 t) w
        Two Fp6 products = 12 dots of 6 terms (a general product needs 12 dots of 12 terms)."""
        assert d is not a
        with self.pool("TMP", "SQR"):
            self._f12_sqr(d, a)

    def _f12_sqr(self, d, a):
        A_ = self.cfg.xi_a
        c0 = [a[0], a[2], a[4]]
        c1 = [a[1], a[3], a[5]]
        S = [self.fp2("SQR.S%d" % j) for j in range(3)]      # c0 + c1
        B = [self.fp2("SQR.B%d" % j) for j in range(3)]      # c0 + v c1 = (c0_0 + xi c1_2, c0_1 + c1_0, c0_2 + c1_1)
        lin = []
        for j in range(3):
            for c in range(2):
                lin.append(self.lin(S[j][c], (c0[j][c], 1), (c1[j][c], 1)))
        # xi * c1_2 = (A x - y, x + A y)
        lin += [self.lin(B[0][0], (c0[0][0], 1), (c1[2][0], A_), (c1[2][1], -1)),
                self.lin(B[0][1], (c0[0][1], 1), (c1[2][0], 1), (c1[2][1], A_))]
        for j in (1, 2):
            for c in range(2):
                lin.append(self.lin(B[j][c], (c0[j][c], 1), (c1[j - 1][c], 1)))
        self.lin_rounds(lin)
        # derived copies of the second operands (c1 for t, B for s)
        d1 = [self.der_slots("SQR.d1_%d" % j) for j in range(3)]
        dB = [self.der_slots("SQR.dB_%d" % j) for j in range(3)]
        lin = []
        for j in range(3):
            lin += self.derived(c1[j], d1[j]) + self.derived(B[j], dB[j])
        self.lin_rounds(lin)

        def fp6_dots(dst, u, v, dv):
            """dst = u * v in Fp6; dv = derived copies of v."""
            out = []
            for k in range(3):
                re, im = [], []
                for i in range(3):
                    j = (k - i) % 3
                    if i <= k:  # no wrap
                        re += [(u[i][0], v[j][0]), (u[i][1], dv[j]["ny"])]
                        im += [(u[i][0], v[j][1]), (u[i][1], v[j][0])]
                    else:       # v^3 = xi
                        re += [(u[i][0], dv[j]["tx"]), (u[i][1], dv[j]["nty"])]
                        im += [(u[i][0], dv[j]["ty"]), (u[i][1], dv[j]["tx"])]
                out += [(dst[k][0], re), (dst[k][1], im)]
            return out
        Tt = [self.fp2("SQR.T%d" % j) for j in range(3)]
        Ss = [self.fp2("SQR.P%d" % j) for j in range(3)]
        self.dot(fp6_dots(Tt, c0, c1, d1) + fp6_dots(Ss, S, B, dB))
        # new c0 = s - t - v t,  v t = (xi t2, t0, t1);  new c1 = 2 t
        nc0 = [d[0], d[2], d[4]]
        nc1 = [d[1], d[3], d[5]]
        lin = [self.lin(nc0[0][0], (Ss[0][0], 1), (Tt[0][0], -1), (Tt[2][0], -A_), (Tt[2][1], 1)),
               self.lin(nc0[0][1], (Ss[0][1], 1), (Tt[0][1], -1), (Tt[2][0], -1), (Tt[2][1], -A_))]
        for j in (1, 2):
            for c in range(2):
                lin.append(self.lin(nc0[j][c], (Ss[j][c], 1), (Tt[j][c], -1), (Tt[j - 1][c], -1)))
        for j in range(3):
            for c in range(2):
                lin.append(self.lin(nc1[j][c], (Tt[j][c], 2)))
        self.lin_rounds(lin)

    def f12_sparse(self, d, a, line):
        """d = a * (sum_c line[pos_c] w^pos_c); line: dict pos -> {x, y, ny, tx, ty, nty} slots."""
        assert d is not a
        tasks = []
        for k in range(6):
            re, im = [], []
            for pos, l in sorted(line.items()):
                i = (k - pos) % 6
                wrapped = k - pos < 0
                if not wrapped:
                    re += [(a[i][0], l["x"]), (a[i][1], l["ny"])]
                    im += [(a[i][0], l["y"]), (a[i][1], l["x"])]
                else:
                    re += [(a[i][0], l["tx"]), (a[i][1], l["nty"])]
                    im += [(a[i][0], l["ty"]), (a[i][1], l["tx"])]
            tasks += [(d[k][0], re), (d[k][1], im)]
        self.dot(tasks)

    def f12_copy(self, d, a):
        self.lin_rounds([self.lin(d[k][c], (a[k][c], 1)) for k in range(6) for c in range(2)])

    def f12_conj(self, d, a):
        """a^(p^6): w -> -w."""
        self.lin_rounds([self.lin(d[k][c], (a[k][c], -1 if k & 1 else 1)) for k in range(6) for c in range(2)])

    def f12_frob(self, d, a, e):
        """d = a^(p^e), e in 1..3: coefficient k -> conj^e(c_k) * gamma_e[k]."""
        assert d is not a
        cfg, p = self.cfg, self.cfg.p
        xi = (cfg.xi_a, 1)
        g1 = f2pow(xi, (p - 1) // 6, p)
        expo = {1: 1, 2: p + 1, 3: p * p + p + 1}[e]
        tasks, lin = [], []
        nys = [self.g("FROB.ny%d" % k) for k in range(6)]
        if e & 1:
            # conj(c) * g = (x g0 + y g1) + (x g1 - y g0) i
            lin = [self.lin(nys[k], (a[k][1], -1)) for k in range(1, 6)]
        for k in range(6):
            x, y = a[k]
            if k == 0:
                continue
            gk = f2pow(f2pow(g1, k, p), expo, p)
            gc = self.const2("GAM%d_%d" % (e, k), gk)
            if e & 1:
                tasks += [(d[k][0], [(x, gc[0]), (y, gc[1])]), (d[k][1], [(x, gc[1]), (nys[k], gc[0])])]
            else:
                assert gk[1] == 0
                tasks += [(d[k][0], [(x, gc[0])]), (d[k][1], [(y, gc[0])])]
        lin += [self.lin(d[0][0], (a[0][0], 1)), self.lin(d[0][1], (a[0][1], -1 if e & 1 else 1))]
        self.lin_rounds(lin)
        self.dot(tasks)

    def f12_cycsqr(self, d, a):
        """Granger-Scott squaring in the cyclotomic subgroup, d distinct from a.  One LIN phase builds scaled
        copies of the operands, one DOT phase produces the outputs directly (so bounds contract):
          (z0', z1') from (z0, z1): z0' = 3(z0^2 + xi z1^2) - 2 z0,  z1' = 6 z0 z1 + 2 z1
          (z4', z5') from (z2, z3): z4' = 3(z2^2 + xi z3^2) - 2 z4,  z5' = 6 z2 z3 + 2 z5
          (z3', z2') from (z4, z5): z3' = 3(z4^2 + xi z5^2) - 2 z3,  z2' = 6 z4 (xi z5) + 2 z2
        tower view: c0 = (a0, a2, a4), c1 = (a1, a3, a5); z0=c0.0 z4=c0.1 z3=c0.2 z2=c1.0 z1=c1.1 z5=c1.2"""
        assert d is not a
        if self.signed:
            return self._f12_cycsqr_signed(d, a)
        z = {0: a[0], 4: a[2], 3: a[4], 2: a[1], 1: a[3], 5: a[5]}
        zd = {0: d[0], 4: d[2], 3: d[4], 2: d[1], 1: d[3], 5: d[5]}
        A = self.cfg.xi_a
        C2, CN2 = self.const("C2", 2), self.const("CN2", -2)
        lin, dots = [], []
        for n, (u, v, e, o, twist_cross) in enumerate(((0, 1, 0, 1, False), (2, 3, 4, 5, False), (4, 5, 3, 2, True))):
            (ux, uy), (vx, vy) = z[u], z[v]
            S = lambda nm: self.g("CYC.%d.%s" % (n, nm))
            u3x, un3y, u6y = S("u3x"), S("un3y"), S("u6y")
            v3tx, v3ty, vn3ty = S("v3tx"), S("v3ty"), S("vn3ty")
            c6x, c6y, cn6y = S("c6x"), S("c6y"), S("cn6y")
            lin += [self.lin(u3x, (ux, 3)), self.lin(un3y, (uy, -3)), self.lin(u6y, (uy, 6)),
                    self.lin(v3tx, (vx, 3 * A), (vy, -3)), self.lin(v3ty, (vx, 3), (vy, 3 * A)), self.lin(vn3ty, (vx, -3), (vy, -3 * A))]
            if not twist_cross:
                lin += [self.lin(c6x, (vx, 6)), self.lin(c6y, (vy, 6)), self.lin(cn6y, (vy, -6))]
            else:  # 6 * xi * v
                lin += [self.lin(c6x, (vx, 6 * A), (vy, -6)), self.lin(c6y, (vx, 6), (vy, 6 * A)), self.lin(cn6y, (vx, -6), (vy, -6 * A))]
            dots += [(zd[e][0], [(ux, u3x), (uy, un3y), (vx, v3tx), (vy, vn3ty), (z[e][0], CN2)]),
                     (zd[e][1], [(ux, u6y), (vx, v3ty), (vy, v3tx), (z[e][1], CN2)]),
                     (zd[o][0], [(ux, c6x), (uy, cn6y), (z[o][0], C2)]),
                     (zd[o][1], [(ux, c6y), (uy, c6x), (z[o][1], C2)])]
        self.lin_rounds(lin)
        self.dot(dots)

    def _f12_cycsqr_signed(self, d, a):
        """The same squaring on a signed slot file: ONE DOT phase of the 18 products of at most two terms
             p1 = ux^2 - uy^2, p2 = ux uy, p3 = vx^2 - vy^2, p4 = vx vy, p5 = ux vx - uy vy, p6 = ux vy + uy vx
        per (u, v) pair plus the 12 scalings q_e = -2 z_e, q_o = 2 z_o (products with constants, so that the bounds
        contract: a LIN term -2 z would grow the bound by 2.5x per squaring), then ONE LIN phase
             z_e' = 3 (u^2 + xi v^2) - 2 z_e = (3 p1 + 3A p3 - 6 p4 + q_e.x,  6 p2 + 3 p3 + 6A p4 + q_e.y)
             z_o' = 6 [xi] u v + 2 z_o
        (no scaled operand copies, no split dots: two phases per squaring instead of three)."""
        z = {0: a[0], 4: a[2], 3: a[4], 2: a[1], 1: a[3], 5: a[5]}
        zd = {0: d[0], 4: d[2], 3: d[4], 2: d[1], 1: d[3], 5: d[5]}
        A = self.cfg.xi_a
        dots, lin, ps = [], [], {}
        C2, CN2 = self.const("C2", 2), self.const("CN2", -2)
        groups = ((0, 1, 0, 1, False), (2, 3, 4, 5, False), (4, 5, 3, 2, True))
        for n, (u, v, e, o, twist_cross) in enumerate(groups):
            (ux, uy), (vx, vy) = z[u], z[v]
            ps[n] = tuple(self.g("CYS.%d.p%d" % (n, k)) for k in range(1, 7)) + tuple(self.g("CYS.%d.q%d" % (n, k)) for k in range(4))
            p1, p2, p3, p4, p5, p6, qex, qey, qox, qoy = ps[n]
            dots += [(p1, [(ux, ux), (uy, uy, -1)]), (p2, [(ux, uy)]), (p3, [(vx, vx), (vy, vy, -1)]), (p4, [(vx, vy)]),
                     (p5, [(ux, vx), (uy, vy, -1)]), (p6, [(ux, vy), (uy, vx)]),
                     (qex, [(z[e][0], CN2)]), (qey, [(z[e][1], CN2)]), (qox, [(z[o][0], C2)]), (qoy, [(z[o][1], C2)])]
        self.dot(dots)
        for n, (u, v, e, o, twist_cross) in enumerate(groups):
            p1, p2, p3, p4, p5, p6, qex, qey, qox, qoy = ps[n]
            lin += [self.lin(zd[e][0], (p1, 3), (p3, 3 * A), (p4, -6), (qex, 1)),
                    self.lin(zd[e][1], (p2, 6), (p3, 3), (p4, 6 * A), (qey, 1))]
            if not twist_cross:
                lin += [self.lin(zd[o][0], (p5, 6), (qox, 1)), self.lin(zd[o][1], (p6, 6), (qoy, 1))]
            else:
                lin += [self.lin(zd[o][0], (p5, 6 * A), (p6, -6), (qox, 1)), self.lin(zd[o][1], (p5, 6), (p6, 6 * A), (qoy, 1))]
        self.lin_rounds(lin)

    def declare(self, slots, bound):
        """Loop invariant: assume `bound` for these slots from here on (checked with check_le)."""
        for sl in slots:
            assert self.ub.get(sl) is None or self.ub[sl] <= bound, ("invariant smaller than current bound", sl)
            self.ub[sl] = bound

    def check_le(self, slots, bound):
        for sl in slots:
            assert self.ub[sl] <= bound, ("loop invariant violated", sl, self.ub[sl] / self.cfg.p)


# ======================================================================================
# programs
# ======================================================================================

def line_slots(gen, name):
    return {k: gen.g("%s.%s" % (name, k)) for k in ("x", "y", "ny", "tx", "ty", "nty")}


def line_lin(gen, l, pos, x_terms, y_terms):
    """LIN tasks producing the copies of one line coefficient the sparse product needs.
    x_terms / y_terms: list of (slot, coef) giving the coefficient's re / im parts."""
    a = gen.cfg.xi_a
    neg = lambda ts: [(s, -c) for s, c in ts]
    mul = lambda ts, k: [(s, c * k) for s, c in ts]
    out = [gen.lin(l["x"], *x_terms), gen.lin(l["y"], *y_terms), gen.lin(l["ny"], *neg(y_terms))]
    if pos > 0:  # wrapped positions exist
        out += [gen.lin(l["tx"], *(mul(x_terms, a) + neg(y_terms))), gen.lin(l["ty"], *(x_terms + mul(y_terms, a))),
                gen.lin(l["nty"], *(neg(x_terms) + mul(neg(y_terms), a)))]
    return out


def build_miller(gen: Gen):
    """Program MILLER: inputs RAW.xP, RAW.yP, RAW.xQ.x/.y, RAW.yQ.x/.y (plain integers < p);
    output Fp12 register FA (Montgomery form) = f_{lambda,Q}(P) up to subfield factors."""
    cfg, p = gen.cfg, gen.cfg.p
    D = cfg.twist == "D"
    EC = True  # both curves carry E = 3b'Z^2 as the product BB8*BE8 (keeps LIN coefficients small)
    gen.begin("MILLER")
    raw = {n: gen.g("RAW." + n) for n in ("xP", "yP", "xQ.x", "xQ.y", "yQ.x", "yQ.y")}
    for s in raw.values():
        gen.set_ub(s, p - 1)
    xP, yP = gen.g("xP"), gen.g("yP")
    xQ, yQ = gen.fp2("xQ"), gen.fp2("yQ")
    # to Montgomery form
    gen.dot([(xP, [(raw["xP"], gen.R2)]), (yP, [(raw["yP"], gen.R2)]), (xQ[0], [(raw["xQ.x"], gen.R2)]),
             (xQ[1], [(raw["xQ.y"], gen.R2)]), (yQ[0], [(raw["yQ.x"], gen.R2)]), (yQ[1], [(raw["yQ.y"], gen.R2)])])
    X, Y, Z = gen.fp2("X"), gen.fp2("Y"), gen.fp2("Z")
    nX1, nY1, nZ1 = gen.g("nX1"), gen.g("nY1"), gen.g("nZ1")
    nxQ1, nyQ1, nyQ0 = gen.g("nxQ1"), gen.g("nyQ1"), gen.g("nyQ0")
    FA, FB = gen.f12("FA"), gen.f12("FB")
    xi = (cfg.xi_a, 1)
    if D:
        b2 = f2mul((cfg.b, 0), f2inv(xi, p), p)
    else:
        b2 = f2mul((cfg.b, 0), xi, p)
    b3 = (3 * b2[0] % p, 3 * b2[1] % p)
    lin = [gen.lin(X[0], (xQ[0], 1)), gen.lin(X[1], (xQ[1], 1)), gen.lin(Y[0], (yQ[0], 1)), gen.lin(Y[1], (yQ[1], 1)),
           gen.lin(Z[0], (gen.ONE, 1)), gen.lin(Z[1], (gen.ZERO, 1)),
           gen.lin(nX1, (xQ[1], -1)), gen.lin(nY1, (yQ[1], -1)), gen.lin(nZ1, (gen.ZERO, 1)),
           gen.lin(nxQ1, (xQ[1], -1)), gen.lin(nyQ1, (yQ[1], -1)), gen.lin(nyQ0, (yQ[0], -1))]
    if EC:
        # E = 3b' Z^2 is carried as the product BB8 * BE8: initially 1 * 3b'
        BB8, BE8, nBE8y = gen.fp2("BB8"), gen.fp2("BE8"), gen.g("nBE8y")
        b3c = gen.const2("B3", b3)
        lin += [gen.lin(BB8[0], (gen.ONE, 1)), gen.lin(BB8[1], (gen.ZERO, 1)), gen.lin(BE8[0], (b3c[0], 1)),
                gen.lin(BE8[1], (b3c[1], 1)), gen.lin(nBE8y, (b3c[1], -1))]
    gen.lin_rounds(lin)
    # f = 1
    gen.lin_rounds([gen.lin(FA[k][c], (gen.ONE if (k == 0 and c == 0) else gen.ZERO, 1)) for k in range(6) for c in range(2)])

    lines = {"y": line_slots(gen, "LY"), "x": line_slots(gen, "LX"), "c": line_slots(gen, "LC")}
    # positions of the three line coefficients in powers of w
    pos = {"y": 0, "x": 1, "c": 3} if D else {"c": 0, "x": 2, "y": 3}
    linemap = {pos[k]: lines[k] for k in ("y", "x", "c")}

    def dbl_point():
        XY, B, X2, YZ, E = gen.fp2("XY"), gen.fp2("B"), gen.fp2("X2"), gen.fp2("YZ"), gen.fp2("E")
        t = gen.mul2(XY, X, Y, nY1) + gen.mul2(B, Y, Y, nY1) + gen.mul2(X2, X, X, nX1) + gen.mul2(YZ, Y, Z, nZ1)
        if EC:
            t += gen.mul2(E, BB8, BE8, nBE8y)
        else:
            Cz = gen.fp2("Cz")
            t += gen.mul2(Cz, Z, Z, nZ1)
        gen.dot(t)
        F, BmF, BpF = gen.fp2("F"), gen.fp2("BmF"), gen.fp2("BpF")
        nBmFy, nBpFy, nEy, nYZy, nBy = gen.g("nBmFy"), gen.g("nBpFy"), gen.g("nEy"), gen.g("nYZy"), gen.g("nBy")
        lin = []
        if not EC:
            # E = 3 b' C = 12 (1+i) C = 12 (Cx - Cy) + 12 (Cx + Cy) i
            lin1 = [gen.lin(E[0], (Cz[0], 12), (Cz[1], -12)), gen.lin(E[1], (Cz[0], 12), (Cz[1], 12))]
            gen.lin_rounds(lin1)
        lin += [gen.lin(BmF[0], (B[0], 1), (E[0], -3)), gen.lin(BmF[1], (B[1], 1), (E[1], -3)), gen.lin(nBmFy, (B[1], -1), (E[1], 3)),
                gen.lin(BpF[0], (B[0], 1), (E[0], 3)), gen.lin(BpF[1], (B[1], 1), (E[1], 3)), gen.lin(nBpFy, (B[1], -1), (E[1], -3)),
                gen.lin(nEy, (E[1], -1)), gen.lin(nYZy, (YZ[1], -1)), gen.lin(nBy, (B[1], -1))]
        # constant line coefficient lc = B - E
        lin += line_lin(gen, lines["c"], pos["c"], [(B[0], 1), (E[0], -1)], [(B[1], 1), (E[1], -1)])
        gen.lin_rounds(lin)
        X3, S, EE, Z3, LYp, LXp = gen.fp2("X3"), gen.fp2("S"), gen.fp2("EE"), gen.fp2("Z3"), gen.fp2("LYp"), gen.fp2("LXp")
        t = gen.mul2(X3, XY, BmF, nBmFy) + gen.mul2(S, BpF, BpF, nBpFy) + gen.mul2(EE, E, E, nEy) + gen.mul2(Z3, B, YZ, nYZy)
        if EC:
            BB, BE = gen.fp2("BB"), gen.fp2("BE")
            t += gen.mul2(BB, B, B, nBy) + gen.mul2(BE, B, E, nEy)
        t += gen.mul2_fp(LYp, YZ, yP) + gen.mul2_fp(LXp, X2, xP)
        gen.dot(t)
        # (4X3, 4Y3, 4Z3) = (2 XY (B-F), (B+F)^2 - 12 E^2, 8 B YZ)
        lin = [gen.lin(X[0], (X3[0], 2)), gen.lin(X[1], (X3[1], 2)), gen.lin(nX1, (X3[1], -2)),
               gen.lin(Y[0], (S[0], 1), (EE[0], -12)), gen.lin(Y[1], (S[1], 1), (EE[1], -12)), gen.lin(nY1, (S[1], -1), (EE[1], 12)),
               gen.lin(Z[0], (Z3[0], 8)), gen.lin(Z[1], (Z3[1], 8)), gen.lin(nZ1, (Z3[1], -8))]
        if EC:
            lin += [gen.lin(BB8[0], (BB[0], 8)), gen.lin(BB8[1], (BB[1], 8)), gen.lin(BE8[0], (BE[0], 8)),
                    gen.lin(BE8[1], (BE[1], 8)), gen.lin(nBE8y, (BE[1], -8))]
        # ly = H yP = 2 YZ yP ; lx = -3 X^2 xP
        lin += line_lin(gen, lines["y"], pos["y"], [(LYp[0], 2)], [(LYp[1], 2)])
        lin += line_lin(gen, lines["x"], pos["x"], [(LXp[0], -3)], [(LXp[1], -3)])
        gen.lin_rounds(lin)

    def add_point(Qx, Qy, nQxy, nQyx, nQyy, Qyy_pos):
        with gen.pool("TMP", "ADD"):
            _add_point(Qx, Qy, nQxy, nQyx, nQyy, Qyy_pos)

    def _add_point(Qx, Qy, nQxy, nQyx, nQyy, Qyy_pos):
        """T += Q', Q' = (Qx, Qy) with helper slots nQxy = -Qx.y, nQyx = -Qy.x, nQyy = -Qy.y, Qyy_pos = Qy.y"""
        yqZ, xqZ = gen.fp2("yqZ"), gen.fp2("xqZ")
        t = gen.mul2(yqZ, Qy, Z, nZ1) + gen.mul2(xqZ, Qx, Z, nZ1)
        if EC:
            # E = 3b' Z^2 of the current point (= BB8 * BE8); after the addition E' = E * (lam^3)^2
            E = gen.fp2("E")
            t += gen.mul2(E, BB8, BE8, nBE8y)
        gen.dot(t)
        TH, LA = gen.fp2("TH"), gen.fp2("LA")
        nTHy, nLAy = gen.g("nTHy"), gen.g("nLAy")
        gen.lin_rounds([gen.lin(TH[0], (Y[0], 1), (yqZ[0], -1)), gen.lin(TH[1], (Y[1], 1), (yqZ[1], -1)), gen.lin(nTHy, (Y[1], -1), (yqZ[1], 1)),
                        gen.lin(LA[0], (X[0], 1), (xqZ[0], -1)), gen.lin(LA[1], (X[1], 1), (xqZ[1], -1)), gen.lin(nLAy, (X[1], -1), (xqZ[1], 1))])
        Cc, Dd, LCp, LYp, LXp = gen.fp2("Cc"), gen.fp2("Dd"), gen.fp2("LCp"), gen.fp2("LYp"), gen.fp2("LXp")
        t = gen.mul2(Cc, TH, TH, nTHy) + gen.mul2(Dd, LA, LA, nLAy)
        # lc = theta xQ - lam yQ
        t += [(LCp[0], [(TH[0], Qx[0]), (TH[1], nQxy), (LA[0], nQyx), (LA[1], Qyy_pos)]),
              (LCp[1], [(TH[0], Qx[1]), (TH[1], Qx[0]), (LA[0], nQyy), (LA[1], nQyx)])]
        t += gen.mul2_fp(LYp, LA, yP) + gen.mul2_fp(LXp, TH, xP)
        gen.dot(t)
        nDy, nCy = gen.g("nDy"), gen.g("nCy")
        lin = [gen.lin(nDy, (Dd[1], -1)), gen.lin(nCy, (Cc[1], -1))]
        lin += line_lin(gen, lines["y"], pos["y"], [(LYp[0], 1)], [(LYp[1], 1)])
        lin += line_lin(gen, lines["x"], pos["x"], [(LXp[0], -1)], [(LXp[1], -1)])
        lin += line_lin(gen, lines["c"], pos["c"], [(LCp[0], 1)], [(LCp[1], 1)])
        gen.lin_rounds(lin)
        Ee, Ff, Gg = gen.fp2("Ee"), gen.fp2("Ff"), gen.fp2("Gg")
        gen.dot(gen.mul2(Ee, LA, Dd, nDy) + gen.mul2(Ff, Z, Cc, nCy) + gen.mul2(Gg, X, Dd, nDy))
        H, GmH = gen.fp2("H"), gen.fp2("GmH")
        nHy, nGmHy, nEey = gen.g("nHy"), gen.g("nGmHy"), gen.g("nEey")
        gen.lin_rounds([gen.lin(H[0], (Ee[0], 1), (Ff[0], 1), (Gg[0], -2)), gen.lin(H[1], (Ee[1], 1), (Ff[1], 1), (Gg[1], -2)),
                        gen.lin(nHy, (Ee[1], -1), (Ff[1], -1), (Gg[1], 2)),
                        gen.lin(GmH[0], (Gg[0], 3), (Ee[0], -1), (Ff[0], -1)), gen.lin(GmH[1], (Gg[1], 3), (Ee[1], -1), (Ff[1], -1)),
                        gen.lin(nGmHy, (Gg[1], -3), (Ee[1], 1), (Ff[1], 1)), gen.lin(nEey, (Ee[1], -1))])
        X3, Tt, EY, Z3 = gen.fp2("X3"), gen.fp2("Tt"), gen.fp2("EY"), gen.fp2("Z3")
        t = gen.mul2(X3, LA, H, nHy) + gen.mul2(Tt, TH, GmH, nGmHy) + gen.mul2(EY, Ee, Y, nY1) + gen.mul2(Z3, Z, Ee, nEey)
        if EC:
            EE2 = gen.fp2("EE2")
            t += gen.mul2(EE2, Ee, Ee, nEey)
        gen.dot(t)
        lin = []
        if EC:
            lin += [gen.lin(BB8[0], (E[0], 1)), gen.lin(BB8[1], (E[1], 1)), gen.lin(BE8[0], (EE2[0], 1)),
                    gen.lin(BE8[1], (EE2[1], 1)), gen.lin(nBE8y, (EE2[1], -1))]
        lin += [gen.lin(X[0], (X3[0], 1)), gen.lin(X[1], (X3[1], 1)), gen.lin(nX1, (X3[1], -1)),
                gen.lin(Y[0], (Tt[0], 1), (EY[0], -1)), gen.lin(Y[1], (Tt[1], 1), (EY[1], -1)), gen.lin(nY1, (Tt[1], -1), (EY[1], 1)),
                gen.lin(Z[0], (Z3[0], 1)), gen.lin(Z[1], (Z3[1], 1)), gen.lin(nZ1, (Z3[1], -1))]
        gen.lin_rounds(lin)

    def sparse_into_FA_from(Fsrc):
        gen.f12_sparse(FA, Fsrc, linemap)

    # ---- loop invariants (bounds assumed at every re-use of a phase; verified after each iteration)
    INV_F, INV_T = 64 * p, 4096 * p
    fslots = [FA[k][c] for k in range(6) for c in range(2)] + [FB[k][c] for k in range(6) for c in range(2)]
    tslots = [X[0], X[1], Y[0], Y[1], Z[0], Z[1], nX1, nY1, nZ1] + ([BB8[0], BB8[1], BE8[0], BE8[1], nBE8y] if EC else [])

    # ---- main loop
    if cfg.loop_naf:
        digits = naf(cfg.loop)
    else:
        digits = [int(c) for c in bin(cfg.loop)[2:]]
    assert digits[0] == 1
    first = True
    for dgt in digits[1:]:
        if first:
            # f == 1: f^2 * line = line.  Build f = line directly through the sparse product with FA = 1 -> FB -> FA
            dbl_point()
            gen.f12_sparse(FB, FA, linemap)
            first = False
            cur = FB
        else:
            gen.f12_sqr(FB, FA)          # FB = FA^2
            dbl_point()
            gen.f12_sparse(FA, FB, linemap)
            cur = FA
        if dgt != 0:
            if dgt == 1:
                add_point(xQ, yQ, nxQ1, nyQ0, nyQ1, yQ[1])
            else:
                add_point(xQ, (nyQ0, nyQ1), nxQ1, yQ[0], yQ[1], nyQ1)
            other = FA if cur is FB else FB
            gen.f12_sparse(other, cur, linemap)
            cur = other
        if cur is FB:
            gen.f12_copy(FA, FB)
            cur = FA
        gen.check_le(fslots[:12], INV_F)
        gen.check_le(tslots, INV_T)
    if D:
        # Q1 = pi(Q), -Q2 = -pi^2(Q):  pi(x, y) = (conj(x) g2, conj(y) g3), g_k = xi^(k(p-1)/6)
        g1 = f2pow(xi, (p - 1) // 6, p)
        g2c = gen.const2("G2", f2pow(g1, 2, p))
        g3c = gen.const2("G3", f2pow(g1, 3, p))
        Q1x, Q1y, Q2x, Q2y = gen.fp2("Q1x"), gen.fp2("Q1y"), gen.fp2("Q2x"), gen.fp2("Q2y")

        def frobq(dx, dy, sx, sy, nsxy, nsyy):
            # conj(s) * g = (s.x g.x + s.y g.y) + (s.x g.y - s.y g.x) i
            gen.dot([(dx[0], [(sx[0], g2c[0]), (sx[1], g2c[1])]), (dx[1], [(sx[0], g2c[1]), (nsxy, g2c[0])]),
                     (dy[0], [(sy[0], g3c[0]), (sy[1], g3c[1])]), (dy[1], [(sy[0], g3c[1]), (nsyy, g3c[0])])])
        frobq(Q1x, Q1y, xQ, yQ, nxQ1, nyQ1)
        nQ1xy, nQ1yx, nQ1yy = gen.g("nQ1xy"), gen.g("nQ1yx"), gen.g("nQ1yy")
        gen.lin_rounds([gen.lin(nQ1xy, (Q1x[1], -1)), gen.lin(nQ1yx, (Q1y[0], -1)), gen.lin(nQ1yy, (Q1y[1], -1))])
        frobq(Q2x, Q2y, Q1x, Q1y, nQ1xy, nQ1yy)
        nQ2xy, nQ2yx, nQ2yy = gen.g("nQ2xy"), gen.g("nQ2yx"), gen.g("nQ2yy")
        gen.lin_rounds([gen.lin(nQ2xy, (Q2x[1], -1)), gen.lin(nQ2yx, (Q2y[0], -1)), gen.lin(nQ2yy, (Q2y[1], -1))])
        add_point(Q1x, Q1y, nQ1xy, nQ1yx, nQ1yy, Q1y[1])
        gen.f12_sparse(FB, FA, linemap)
        # -Q2 = (Q2x, -Q2y)
        add_point(Q2x, (nQ2yx, nQ2yy), nQ2xy, Q2y[0], Q2y[1], nQ2yy)
        gen.f12_sparse(FA, FB, linemap)
    else:
        gen.f12_conj(FA, FA)  # x < 0
    return raw, FA


def build_miller_p(gen: Gen):
    """Program MILLER of the pipelined slot file "P": one pairing per 32-lane group (one warp), signed DOT terms.

    Two independent dependency chains run through the loop: the point chain (T <- 2T [+ Q], line coefficients)
    and the f chain (f <- f^2 * line).  Their stages are co-issued in the same phases, and the first DOT of
    the next doubling rides with the sparse product of the current one, so a doubling iteration is 4 phases
    (2 DOT of <= 6 terms + 2 LIN) instead of 12:
        B: LIN  sqr operands (S, B, xi-copies)        | BmF, BpF, line c
        C: DOT  t = c0 c1, s = S B                    | X3, S, EE, Z3, BB, BE, ly', lx'
        D: LIN  f^2 = (s - t - v t) + 2 t w           | new X, Y, Z, BB8, BE8, lines y, x
        A: DOT  f = f^2 * line                        | XY, B, X2, YZ, E of the next doubling
    Signed terms remove every negated copy (-y, -xi y) of the 16-lane program.  Same inputs/outputs as build_miller."""
    cfg, p = gen.cfg, gen.cfg.p
    assert gen.signed and gen.lanes == 32
    D = cfg.twist == "D"
    A_ = cfg.xi_a
    # Karatsuba-lane form of the two big DOT phases (3 terms instead of 6): the operand sums have 29 bits, which the
    # signed 64-bit columns only hold for the 10-limb curve
    KD = gen.kdot_fits()
    gen.begin("MILLER")
    with gen.pool("QTMP", "RAW"):   # the raw inputs are dead after the first phase
        raw = {n: gen.g("RAW." + n) for n in ("xP", "yP", "xQ.x", "xQ.y", "yQ.x", "yQ.y")}
    for sl in raw.values():
        gen.set_ub(sl, p - 1)
    xP, yP = gen.g("xP"), gen.g("yP")
    xQ, yQ, nyQ = gen.fp2("xQ"), gen.fp2("yQ"), gen.fp2("nyQ")
    gen.dot([(xP, [(raw["xP"], gen.R2)]), (yP, [(raw["yP"], gen.R2)]), (xQ[0], [(raw["xQ.x"], gen.R2)]),
             (xQ[1], [(raw["xQ.y"], gen.R2)]), (yQ[0], [(raw["yQ.x"], gen.R2)]), (yQ[1], [(raw["yQ.y"], gen.R2)])])
    X, Y, Z = gen.fp2("X"), gen.fp2("Y"), gen.fp2("Z")
    BB8, BE8 = gen.fp2("BB8"), gen.fp2("BE8")
    FA, FB = gen.f12("FA"), gen.f12("FB")
    xi = (A_, 1)
    b2 = f2mul((cfg.b, 0), f2inv(xi, p), p) if D else f2mul((cfg.b, 0), xi, p)
    b3c = gen.const2("B3", (3 * b2[0] % p, 3 * b2[1] % p))
    lin = [gen.lin(X[0], (xQ[0], 1)), gen.lin(X[1], (xQ[1], 1)), gen.lin(Y[0], (yQ[0], 1)), gen.lin(Y[1], (yQ[1], 1)),
           gen.lin(Z[0], (gen.ONE, 1)), gen.lin(Z[1], (gen.ZERO, 1)),
           gen.lin(nyQ[0], (yQ[0], -1)), gen.lin(nyQ[1], (yQ[1], -1)),
           # E = 3b' Z^2 is carried as the product BB8 * BE8: initially 1 * 3b'
           gen.lin(BB8[0], (gen.ONE, 1)), gen.lin(BB8[1], (gen.ZERO, 1)), gen.lin(BE8[0], (b3c[0], 1)), gen.lin(BE8[1], (b3c[1], 1))]
    lin += [gen.lin(FA[k][c], (gen.ONE if (k == 0 and c == 0) else gen.ZERO, 1)) for k in range(6) for c in range(2)]
    gen.lin_rounds(lin)

    pos = {"y": 0, "x": 1, "c": 3} if D else {"c": 0, "x": 2, "y": 3}

    def lslots(name, ps):
        return {k: gen.g("%s.%s" % (name, k)) for k in (("x", "y", "tx", "ty") if ps > 0 else ("x", "y"))}
    lines = {"y": lslots("LY", pos["y"]), "x": lslots("LX", pos["x"]), "c": lslots("LC", pos["c"])}
    linemap = {pos[k]: lines[k] for k in ("y", "x", "c")}

    def m2(dst, u, v):       # Fp2 product
        return [(dst[0], [(u[0], v[0]), (u[1], v[1], -1)]), (dst[1], [(u[0], v[1]), (u[1], v[0])])]

    def m2fp(dst, u, s_):
        return [(dst[0], [(u[0], s_)]), (dst[1], [(u[1], s_)])]

    def line_tasks(l, ps, x_terms, y_terms):
        neg = lambda ts: [(s_, -c) for s_, c in ts]
        mul = lambda ts, k: [(s_, c * k) for s_, c in ts]
        out = [gen.lin(l["x"], *x_terms), gen.lin(l["y"], *y_terms)]
        if ps > 0:   # positions that wrap around w^6 = xi need xi * coefficient
            out += [gen.lin(l["tx"], *(mul(x_terms, A_) + neg(y_terms))), gen.lin(l["ty"], *(x_terms + mul(y_terms, A_)))]
        return out

    def sparse_triples(d, a):
        out = []
        for k in range(6):
            terms = []
            for ps, l in sorted(linemap.items()):
                i = (k - ps) % 6
                terms.append((a[i], (l["x"], l["y"]) if k - ps >= 0 else (l["tx"], l["ty"])))
            out.append((d[k][0], d[k][1], terms))
        return out

    def sparse_tasks(d, a):
        tasks = []
        for k in range(6):
            re, im = [], []
            for ps, l in sorted(linemap.items()):
                i = (k - ps) % 6
                if k - ps >= 0:
                    re += [(a[i][0], l["x"]), (a[i][1], l["y"], -1)]
                    im += [(a[i][0], l["y"]), (a[i][1], l["x"])]
                else:
                    re += [(a[i][0], l["tx"]), (a[i][1], l["ty"], -1)]
                    im += [(a[i][0], l["ty"]), (a[i][1], l["tx"])]
            tasks += [(d[k][0], re), (d[k][1], im)]
        return tasks

    # ---- f^2 by complex squaring over Fp6 = Fp2[v]/(v^3 - xi), v = w^2 (see Gen._f12_sqr), in three stages
    # The temporaries of the squaring (live from phase B to phase D of a doubling), of the addition step (live only
    # between two doublings) and of the product-tree multiplication (after the loop) are never live together: they
    # share physical slots (pool "PTMP"), which keeps the per-warp slot file small enough for three blocks per SM.
    with gen.pool("PTMP", "SQR"):
        S = [gen.fp2("SQ.S%d" % j) for j in range(3)]
        Bq = [gen.fp2("SQ.B%d" % j) for j in range(3)]
        d1 = {j: {"tx": gen.g("SQ.d1_%d.tx" % j), "ty": gen.g("SQ.d1_%d.ty" % j)} for j in (1, 2)}
        dB = {j: {"tx": gen.g("SQ.dB_%d.tx" % j), "ty": gen.g("SQ.dB_%d.ty" % j)} for j in (1, 2)}
        Tt = [gen.fp2("SQ.T%d" % j) for j in range(3)]
        Ss = [gen.fp2("SQ.P%d" % j) for j in range(3)]

    def sqr_lin_a(a):
        c0, c1 = [a[0], a[2], a[4]], [a[1], a[3], a[5]]
        t = []
        for j in range(3):
            for c in range(2):
                t.append(gen.lin(S[j][c], (c0[j][c], 1), (c1[j][c], 1)))
        t += [gen.lin(Bq[0][0], (c0[0][0], 1), (c1[2][0], A_), (c1[2][1], -1)),
              gen.lin(Bq[0][1], (c0[0][1], 1), (c1[2][0], 1), (c1[2][1], A_))]
        for j in (1, 2):
            for c in range(2):
                t.append(gen.lin(Bq[j][c], (c0[j][c], 1), (c1[j - 1][c], 1)))
        for j in (1, 2):   # xi-copies of the second operands (only v_1, v_2 ever wrap)
            t += [gen.lin(d1[j]["tx"], (c1[j][0], A_), (c1[j][1], -1)), gen.lin(d1[j]["ty"], (c1[j][0], 1), (c1[j][1], A_)),
                  gen.lin(dB[j]["tx"], (c0[j][0], A_), (c1[j - 1][0], A_), (c0[j][1], -1), (c1[j - 1][1], -1)),
                  gen.lin(dB[j]["ty"], (c0[j][0], 1), (c1[j - 1][0], 1), (c0[j][1], A_), (c1[j - 1][1], A_))]
        return t

    def fp6_dots(dst, u, v, dv):
        out = []
        for k in range(3):
            re, im = [], []
            for i in range(3):
                j = (k - i) % 3
                if i <= k:
                    re += [(u[i][0], v[j][0]), (u[i][1], v[j][1], -1)]
                    im += [(u[i][0], v[j][1]), (u[i][1], v[j][0])]
                else:       # v^3 = xi
                    re += [(u[i][0], dv[j]["tx"]), (u[i][1], dv[j]["ty"], -1)]
                    im += [(u[i][0], dv[j]["ty"]), (u[i][1], dv[j]["tx"])]
            out += [(dst[k][0], re), (dst[k][1], im)]
        return out

    def sqr_dot(a):
        c0, c1 = [a[0], a[2], a[4]], [a[1], a[3], a[5]]
        return fp6_dots(Tt, c0, c1, d1) + fp6_dots(Ss, S, Bq, dB)

    def fp6_triples(dst, u, v, dv):
        out = []
        for k in range(3):
            terms = []
            for i in range(3):
                j = (k - i) % 3
                terms.append((u[i], v[j] if i <= k else (dv[j]["tx"], dv[j]["ty"])))
            out.append((dst[k][0], dst[k][1], terms))
        return out

    def sqr_triples(a):
        c0, c1 = [a[0], a[2], a[4]], [a[1], a[3], a[5]]
        return fp6_triples(Tt, c0, c1, d1) + fp6_triples(Ss, S, Bq, dB)

    def big_dot(triples, schoolbook, plain):
        """One of the two wide DOT phases of an iteration: Karatsuba lanes when the curve allows it."""
        if KD:
            gen.kdot(triples, plain)
        else:
            gen.dot(schoolbook + plain)

    def sqr_lin_b(d):
        nc0, nc1 = [d[0], d[2], d[4]], [d[1], d[3], d[5]]
        t = [gen.lin(nc0[0][0], (Ss[0][0], 1), (Tt[0][0], -1), (Tt[2][0], -A_), (Tt[2][1], 1)),
             gen.lin(nc0[0][1], (Ss[0][1], 1), (Tt[0][1], -1), (Tt[2][0], -1), (Tt[2][1], -A_))]
        for j in (1, 2):
            for c in range(2):
                t.append(gen.lin(nc0[j][c], (Ss[j][c], 1), (Tt[j][c], -1), (Tt[j - 1][c], -1)))
        for j in range(3):
            for c in range(2):
                t.append(gen.lin(nc1[j][c], (Tt[j][c], 2)))
        return t

    # ---- doubling: homogeneous projective, E = 3b'Z^2 = BB8 * BE8
    # doubling temporaries are dead during an addition step and vice versa: second pool "QTMP"
    E, LYp, LXp = gen.fp2("E"), gen.fp2("LYp"), gen.fp2("LXp")   # used by both steps
    with gen.pool("QTMP", "DBL"):
        XY, Bd, X2, YZ = gen.fp2("XY"), gen.fp2("B"), gen.fp2("X2"), gen.fp2("YZ")
        BmF, BpF = gen.fp2("BmF"), gen.fp2("BpF")
        X3, EE, Z3, BB, BE = gen.fp2("X3"), gen.fp2("EE"), gen.fp2("Z3"), gen.fp2("BB"), gen.fp2("BE")

    def dbl_dot1():
        return m2(XY, X, Y) + m2(Bd, Y, Y) + m2(X2, X, X) + m2(YZ, Y, Z) + m2(E, BB8, BE8)

    def dbl_lin1():
        t = [gen.lin(BmF[0], (Bd[0], 1), (E[0], -3)), gen.lin(BmF[1], (Bd[1], 1), (E[1], -3)),
             gen.lin(BpF[0], (Bd[0], 1), (E[0], 3)), gen.lin(BpF[1], (Bd[1], 1), (E[1], 3))]
        return t + line_tasks(lines["c"], pos["c"], [(Bd[0], 1), (E[0], -1)], [(Bd[1], 1), (E[1], -1)])

    def dbl_dot2():
        # (B + 3E)^2 = B^2 + 6 B E + 9 E^2 is a combination of BB, BE, EE: no product of its own (14 tasks, so that the
        # 18 Karatsuba lanes of the squaring fit beside them)
        return (m2(X3, XY, BmF) + m2(EE, E, E) + m2(Z3, Bd, YZ) + m2(BB, Bd, Bd) + m2(BE, Bd, E) +
                m2fp(LYp, YZ, yP) + m2fp(LXp, X2, xP))

    def dbl_lin2():
        # (4X3, 4Y3, 4Z3) = (2 XY (B-F), (B+F)^2 - 12 E^2, 8 B YZ), F = 3E: (B+F)^2 - 12 E^2 = BB + 6 BE - 3 EE;
        # ly = 2 YZ yP,  lx = -3 X^2 xP
        t = [gen.lin(X[0], (X3[0], 2)), gen.lin(X[1], (X3[1], 2)),
             gen.lin(Y[0], (BB[0], 1), (BE[0], 6), (EE[0], -3)), gen.lin(Y[1], (BB[1], 1), (BE[1], 6), (EE[1], -3)),
             gen.lin(Z[0], (Z3[0], 8)), gen.lin(Z[1], (Z3[1], 8)),
             gen.lin(BB8[0], (BB[0], 8)), gen.lin(BB8[1], (BB[1], 8)), gen.lin(BE8[0], (BE[0], 8)), gen.lin(BE8[1], (BE[1], 8))]
        t += line_tasks(lines["y"], pos["y"], [(LYp[0], 2)], [(LYp[1], 2)])
        t += line_tasks(lines["x"], pos["x"], [(LXp[0], -3)], [(LXp[1], -3)])
        return t

    # ---- mixed addition T += Q' = (Qx, Qy) with its line, four DOT + four LIN stages
    with gen.pool("QTMP", "ADD"):      # 11 Fp2 = as many slots as the doubling temporaries
        yqZ, xqZ, TH, LA = gen.fp2("yqZ"), gen.fp2("xqZ"), gen.fp2("TH"), gen.fp2("LA")
        Cc, Dd, LCp = gen.fp2("Cc"), gen.fp2("Dd"), gen.fp2("LCp")
        Ee, Ff, Gg, H = gen.fp2("Ee"), gen.fp2("Ff"), gen.fp2("Gg"), gen.fp2("H")
    with gen.pool("PTMP", "ADD"):      # the rest next to the squaring temporaries (dead between two doublings)
        GmH = gen.fp2("GmH")
        AX3, ATt, EY, AZ3, EE2 = gen.fp2("AX3"), gen.fp2("ATt"), gen.fp2("EY"), gen.fp2("AZ3"), gen.fp2("EE2")
        if D:
            Q1x, Q1y, Q2x, Q2y, nQ2y = gen.fp2("Q1x"), gen.fp2("Q1y"), gen.fp2("Q2x"), gen.fp2("Q2y"), gen.fp2("nQ2y")

    def add_dot1(Qx, Qy):
        return m2(yqZ, Qy, Z) + m2(xqZ, Qx, Z) + m2(E, BB8, BE8)

    def add_lin1():
        return [gen.lin(TH[0], (Y[0], 1), (yqZ[0], -1)), gen.lin(TH[1], (Y[1], 1), (yqZ[1], -1)),
                gen.lin(LA[0], (X[0], 1), (xqZ[0], -1)), gen.lin(LA[1], (X[1], 1), (xqZ[1], -1))]

    def add_dot2(Qx, Qy):
        # lc = theta xQ - lam yQ
        t = m2(Cc, TH, TH) + m2(Dd, LA, LA)
        t += [(LCp[0], [(TH[0], Qx[0]), (TH[1], Qx[1], -1), (LA[0], Qy[0], -1), (LA[1], Qy[1])]),
              (LCp[1], [(TH[0], Qx[1]), (TH[1], Qx[0]), (LA[0], Qy[1], -1), (LA[1], Qy[0], -1)])]
        return t + m2fp(LYp, LA, yP) + m2fp(LXp, TH, xP)

    def add_dot2_phase(Qx, Qy):
        """theta^2, lam^2 and lc = theta xQ - lam yQ as Karatsuba triples (2 terms) when the curve allows it: the 4-term
        schoolbook form of lc would make the whole phase 4 terms deep."""
        if KD:
            gen.kdot([(Cc[0], Cc[1], [(TH, TH)]), (Dd[0], Dd[1], [(LA, LA)]), (LCp[0], LCp[1], [(TH, Qx), (LA, Qy, -1)])],
                     m2fp(LYp, LA, yP) + m2fp(LXp, TH, xP))
        else:
            gen.dot(add_dot2(Qx, Qy))

    def add_lin2():
        return (line_tasks(lines["y"], pos["y"], [(LYp[0], 1)], [(LYp[1], 1)]) +
                line_tasks(lines["x"], pos["x"], [(LXp[0], -1)], [(LXp[1], -1)]) +
                line_tasks(lines["c"], pos["c"], [(LCp[0], 1)], [(LCp[1], 1)]))

    def add_dot3():
        return m2(Ee, LA, Dd) + m2(Ff, Z, Cc) + m2(Gg, X, Dd)

    def add_lin3():
        return [gen.lin(H[0], (Ee[0], 1), (Ff[0], 1), (Gg[0], -2)), gen.lin(H[1], (Ee[1], 1), (Ff[1], 1), (Gg[1], -2)),
                gen.lin(GmH[0], (Gg[0], 3), (Ee[0], -1), (Ff[0], -1)), gen.lin(GmH[1], (Gg[1], 3), (Ee[1], -1), (Ff[1], -1))]

    def add_dot4():
        return m2(AX3, LA, H) + m2(ATt, TH, GmH) + m2(EY, Ee, Y) + m2(AZ3, Z, Ee) + m2(EE2, Ee, Ee)

    def add_lin4():
        # E' = E * (lam^3)^2: carried as BB8 = E, BE8 = Ee^2
        return [gen.lin(BB8[0], (E[0], 1)), gen.lin(BB8[1], (E[1], 1)), gen.lin(BE8[0], (EE2[0], 1)), gen.lin(BE8[1], (EE2[1], 1)),
                gen.lin(X[0], (AX3[0], 1)), gen.lin(X[1], (AX3[1], 1)),
                gen.lin(Y[0], (ATt[0], 1), (EY[0], -1)), gen.lin(Y[1], (ATt[1], 1), (EY[1], -1)),
                gen.lin(Z[0], (AZ3[0], 1)), gen.lin(Z[1], (AZ3[1], 1))]

    def lin1(tasks):
        assert len(tasks) <= gen.lanes, len(tasks)
        gen.lin_rounds(tasks)

    def add_step(Qx, Qy, more):
        """FB = f^2, dbl line ready.  Ends with f in FA and (if `more`) the first DOT of the next doubling issued."""
        big_dot(sparse_triples(FA, FB), sparse_tasks(FA, FB), add_dot1(Qx, Qy))
        lin1(add_lin1())
        add_dot2_phase(Qx, Qy)
        lin1(add_lin2())
        big_dot(sparse_triples(FB, FA), sparse_tasks(FB, FA), add_dot3())
        lin1(add_lin3() + [gen.lin(FA[k][c], (FB[k][c], 1)) for k in range(6) for c in range(2)])
        gen.dot(add_dot4())
        lin1(add_lin4())
        if more:
            gen.dot(dbl_dot1())

    INV_F, INV_T = 64 * p, 4096 * p
    fslots = [FA[k][c] for k in range(6) for c in range(2)]
    tslots = [X[0], X[1], Y[0], Y[1], Z[0], Z[1], BB8[0], BB8[1], BE8[0], BE8[1]]
    digits = naf(cfg.loop) if cfg.loop_naf else [int(c) for c in bin(cfg.loop)[2:]]
    assert digits[0] == 1
    body = digits[1:]
    gen.dot(dbl_dot1())
    for n_, dgt in enumerate(body):
        last = n_ == len(body) - 1
        more = (not last) or False
        lin1(sqr_lin_a(FA) + dbl_lin1())
        big_dot(sqr_triples(FA), sqr_dot(FA), dbl_dot2())
        lin1(sqr_lin_b(FB) + dbl_lin2())
        if dgt == 0:
            big_dot(sparse_triples(FA, FB), sparse_tasks(FA, FB), dbl_dot1() if more else [])
        elif dgt == 1:
            add_step(xQ, yQ, more)
        else:
            add_step(xQ, nyQ, more)
        gen.check_le(fslots, INV_F)
        gen.check_le(tslots, INV_T)
    if D:
        # Q1 = pi(Q), -Q2 = -pi^2(Q):  pi(x, y) = (conj(x) g2, conj(y) g3), g_k = xi^(k(p-1)/6)
        g1 = f2pow(xi, (p - 1) // 6, p)
        g2c = gen.const2("G2", f2pow(g1, 2, p))
        g3c = gen.const2("G3", f2pow(g1, 3, p))

        def frobq(dx, dy, sx, sy):
            # conj(s) * g = (s.x g.x + s.y g.y) + (s.x g.y - s.y g.x) i
            return [(dx[0], [(sx[0], g2c[0]), (sx[1], g2c[1])]), (dx[1], [(sx[0], g2c[1]), (sx[1], g2c[0], -1)]),
                    (dy[0], [(sy[0], g3c[0]), (sy[1], g3c[1])]), (dy[1], [(sy[0], g3c[1]), (sy[1], g3c[0], -1)])]
        gen.dot(frobq(Q1x, Q1y, xQ, yQ))
        gen.dot(frobq(Q2x, Q2y, Q1x, Q1y))
        lin1([gen.lin(nQ2y[0], (Q2y[0], -1)), gen.lin(nQ2y[1], (Q2y[1], -1))])

        def tail_add(Qx, Qy, src, dst):
            gen.dot(add_dot1(Qx, Qy))
            lin1(add_lin1())
            add_dot2_phase(Qx, Qy)
            lin1(add_lin2())
            big_dot(sparse_triples(dst, src), sparse_tasks(dst, src), add_dot3())
            lin1(add_lin3())
            gen.dot(add_dot4())
            lin1(add_lin4())
        tail_add(Q1x, Q1y, FA, FB)
        # the last addition only contributes its line: the point update is not needed
        gen.dot(add_dot1(Q2x, nQ2y))
        lin1(add_lin1())
        add_dot2_phase(Q2x, nQ2y)
        lin1(add_lin2())
        big_dot(sparse_triples(FA, FB), sparse_tasks(FA, FB), [])
    else:
        gen.f12_conj(FA, FA)  # x < 0
    return raw, FA


TREE_P = 48   # invariant of the P-file product tree: every operand and product is below TREE_P * p


def build_mulacc_p(gen: Gen, FA):
    """Program MULACC of the signed 32-lane file: FA <- FA * GB in place (product trees).  Every dot of 12 terms is
    computed as two half dots on two lanes (temporaries), then one LIN phase sums the halves into FA."""
    p, A_ = gen.cfg.p, gen.cfg.xi_a
    with gen.pool("PTMP", "MUL"):
        GB = gen.f12("GB")
    gen.begin("MULACC")
    for k in range(6):
        for c in range(2):
            gen.set_ub(FA[k][c], TREE_P * p)
            gen.set_ub(GB[k][c], TREE_P * p)
    with gen.pool("PTMP", "MUL"):
        dv = {j: {"tx": gen.g("MUL.d%d.tx" % j), "ty": gen.g("MUL.d%d.ty" % j)} for j in range(1, 6)}   # xi * GB_j (j = 0 never wraps)
        mul_tmp = {(k, c, h): gen.g("MUL.%s%d_%d" % (h, k, c)) for k in range(6) for c in range(2) for h in ("lo", "hi")}
    lin = []
    for j in range(1, 6):
        lin += [gen.lin(dv[j]["tx"], (GB[j][0], A_), (GB[j][1], -1)), gen.lin(dv[j]["ty"], (GB[j][0], 1), (GB[j][1], A_))]
    gen.lin_rounds(lin)
    halves, comb = [], []
    for k in range(6):
        re, im = [], []
        for i in range(6):
            j = (k - i) % 6
            if i <= k:
                re.append([(FA[i][0], GB[j][0]), (FA[i][1], GB[j][1], -1)])
                im.append([(FA[i][0], GB[j][1]), (FA[i][1], GB[j][0])])
            else:
                re.append([(FA[i][0], dv[j]["tx"]), (FA[i][1], dv[j]["ty"], -1)])
                im.append([(FA[i][0], dv[j]["ty"]), (FA[i][1], dv[j]["tx"])])
        for c, parts in ((0, re), (1, im)):
            lo, hi = mul_tmp[(k, c, "lo")], mul_tmp[(k, c, "hi")]
            halves += [(lo, sum(parts[:3], [])), (hi, sum(parts[3:], []))]
            comb.append(gen.lin(FA[k][c], (lo, 1), (hi, 1)))
    gen.dot(halves)
    gen.lin_rounds(comb)
    return GB


def build_mulacc(gen: Gen):
    """Programs MUL_AB (FB = FA * GB) and MUL_BA (FA = FB * GB) for product trees."""
    p = gen.cfg.p
    FA, FB, GB = gen.f12("FA"), gen.f12("FB"), gen.f12("GB")
    for name, d, a in (("MUL_AB", FB, FA), ("MUL_BA", FA, FB)):
        gen.begin(name)
        for k in range(6):
            for c in range(2):
                gen.set_ub(a[k][c], 64 * p)
                gen.set_ub(GB[k][c], 64 * p)
        gen.f12_mul(d, a, GB)
    return FA, FB, GB


def build_final_exp(gen: Gen):
    """Program FINALEXP: FA <- FA^((p^12-1)/r), then OUT <- canonical-domain value (times RAW1)."""
    cfg, p = gen.cfg, gen.cfg.p
    gen.begin("FINALEXP")
    regs = {n: gen.f12(n) for n in ("FA", "FB", "R2", "R3", "R4", "R5", "R6", "R7", "R8", "R9")}
    FA = regs["FA"]
    for k in range(6):
        for c in range(2):
            gen.set_ub(FA[k][c], 64 * p)
    M, T0, T1 = regs["FB"], regs["R2"], regs["R3"]

    # ---- inversion of FA -> T1.  N = FA * conj(FA) lies in Fp6 (even coefficients only)
    gen.f12_conj(T0, FA)
    N = regs["R4"]
    gen.f12_mul(N, FA, T0)
    d0, d1, d2 = N[0], N[2], N[4]   # Fp6 = Fp2[v], v = w^2
    dd = [gen.der_slots("INV.d%d" % j) for j in range(3)]
    lt = []
    for j, dj in enumerate((d0, d1, d2)):
        lt += gen.derived(dj, dd[j])
    gen.lin_rounds(lt)
    t0, t1, t2 = gen.fp2("INV.t0"), gen.fp2("INV.t1"), gen.fp2("INV.t2")
    # t0 = d0^2 - xi d1 d2 ; t1 = xi d2^2 - d0 d1 ; t2 = d1^2 - d0 d2
    # products with a negated / xi-twisted second operand use the derived copies:
    #   u * v        : re = ux vx + uy (-vy),  im = ux vy + uy vx
    #   u * (xi v)   : re = ux tvx + uy ntvy,  im = ux tvy + uy tvx
    #   -(u * v)     : re = ux (-vx) + uy vy,  im = ux (-vy) + uy (-vx)        needs -vx: extra LIN below
    nd = [gen.g("INV.ndx%d" % j) for j in range(3)]   # -d_j.x
    ntx = [gen.g("INV.ntx%d" % j) for j in range(3)]  # -(xi d_j).x
    gen.lin_rounds([gen.lin(nd[j], ((d0, d1, d2)[j][0], -1)) for j in range(3)] +
                   [gen.lin(ntx[j], (dd[j]["tx"], -1)) for j in range(3)])
    D_ = (d0, d1, d2)

    def prod(u, j, mode):
        """terms of u * op(d_j): mode '+': d_j, 'x': xi d_j, '-': -d_j, '-x': -xi d_j"""
        v, dv = D_[j], dd[j]
        if mode == "+":
            return [(u[0], v[0]), (u[1], dv["ny"])], [(u[0], v[1]), (u[1], v[0])]
        if mode == "x":
            return [(u[0], dv["tx"]), (u[1], dv["nty"])], [(u[0], dv["ty"]), (u[1], dv["tx"])]
        if mode == "-":
            return [(u[0], nd[j]), (u[1], v[1])], [(u[0], dv["ny"]), (u[1], nd[j])]
        if mode == "-x":
            return [(u[0], ntx[j]), (u[1], dv["ty"])], [(u[0], dv["nty"]), (u[1], ntx[j])]
        raise ValueError(mode)

    def comb(dst, parts):
        re, im = [], []
        for u, j, mode in parts:
            r_, i_ = prod(u, j, mode)
            re += r_
            im += i_
        return [(dst[0], re), (dst[1], im)]
    gen.dot(comb(t0, [(d0, 0, "+"), (d1, 2, "-x")]) + comb(t1, [(d2, 2, "x"), (d0, 1, "-")]) + comb(t2, [(d1, 1, "+"), (d0, 2, "-")]))
    # n = d0 t0 + xi (d2 t1 + d1 t2)
    n = gen.fp2("INV.n")
    gen.dot(comb(n, [(t0, 0, "+"), (t1, 2, "x"), (t2, 1, "x")]))
    # Fp2 inverse of n: nn = n.x^2 + n.y^2 ; n^-1 = (n.x, -n.y) / nn
    nn = gen.g("INV.nn")
    gen.dot([(nn, [(n[0], n[0]), (n[1], n[1])])])
    # Fp inverse of the norm: one single-lane binary-GCD phase (a Fermat chain would be ~380 dependent DOT phases)
    inv = gen.g("INV.inv")
    gen.inv(inv, nn)
    ni = gen.fp2("INV.ni")
    nny = gen.g("INV.nny")
    gen.lin_rounds([gen.lin(nny, (n[1], -1))])
    gen.dot([(ni[0], [(n[0], inv)]), (ni[1], [(nny, inv)])])
    nniy = gen.g("INV.nniy")
    gen.lin_rounds([gen.lin(nniy, (ni[1], -1))])
    # D^-1 = (t0, t1, t2) * n^-1 placed in an Fp12 register (odd coefficients zero), then T1 = conj(FA) * D^-1
    DI = regs["R5"]
    tasks = []
    for j, tj in enumerate((t0, t1, t2)):
        tasks += gen.mul2(DI[2 * j], tj, ni, nniy)
    gen.dot(tasks)
    gen.lin_rounds([gen.lin(DI[k][c], (gen.ZERO, 1)) for k in (1, 3, 5) for c in range(2)])
    gen.f12_mul(T1, T0, DI)            # T1 = FA^-1
    # ---- easy part: m = (conj(FA) * FA^-1)^(p^2+1)
    E1 = regs["R4"]
    gen.f12_mul(E1, T0, T1)
    gen.f12_frob(T0, E1, 2)
    gen.f12_mul(M, T0, E1)             # M = m

    # ---- hard part, written against a tiny Fp12 register allocator
    pool = [regs[n] for n in ("R2", "R3", "R4", "R5", "R6", "R7", "R8", "R9")] + [gen.f12("R%d" % i) for i in range(10, 16)]
    live = [M]

    def alloc():
        r_ = pool.pop()
        live.append(r_)
        return r_

    def free(*rs):
        for r_ in rs:
            if r_ is M:
                continue
            live.remove(r_)
            pool.append(r_)

    def mul(a, b):
        d = alloc(); gen.f12_mul(d, a, b); return d

    def csq(a):
        d = alloc(); gen.f12_cycsqr(d, a); return d

    def frob(a, e):
        d = alloc(); gen.f12_frob(d, a, e); return d

    def conj(a):
        d = alloc(); gen.f12_conj(d, a); return d

    def wnaf3(e):
        """Signed digits of e in {0, +-1, +-3} (width-3 non-adjacent form), least significant first."""
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

    def cyc_pow(base, e):
        """base^e for base in the cyclotomic subgroup (fresh register returned): cyclotomic squarings and a signed
        width-3 window -- the table is base^{+-1}, base^{+-3}, the inverses being conjugates -- so a 63-bit exponent
        costs ~16 multiplications instead of one per set bit."""
        digs = wnaf3(e)
        assert sum(d << i for i, d in enumerate(digs)) == e
        if bin(e).count("1") - sum(1 for d in digs if d) < 4:
            digs = [int(ch) for ch in bin(e)[:1:-1]]     # sparse exponent (bls12-381 x): the table would cost more than it saves
        tbl, owned = {1: base}, []
        if any(abs(d) == 3 for d in digs):
            sq = csq(base)
            tbl[3] = mul(sq, base)
            free(sq)
            owned.append(tbl[3])
        for k in (1, 3):
            if k in tbl and any(d == -k for d in digs):
                tbl[-k] = conj(tbl[k])
                owned.append(tbl[-k])
        der = {k: gen.f12_derived(v, "PDER%s%d" % ("m" if k < 0 else "p", abs(k))) for k, v in sorted(tbl.items())}
        top = len(digs) - 1
        cur_, fresh = tbl[digs[top]], False      # the leading digit of a width-3 NAF of a positive number is positive
        assert digs[top] > 0
        for i in range(top - 1, -1, -1):
            nx = csq(cur_)
            if fresh:
                free(cur_)
            cur_, fresh = nx, True
            if digs[i]:
                nx = alloc()
                gen.f12_mul(nx, cur_, tbl[digs[i]], bder=der[digs[i]])
                free(cur_)
                cur_ = nx
        assert fresh, "exponents of the final exponentiation have more than one digit"
        free(*owned)
        return cur_

    if cfg.name == "BN254":
        u = cfg.u
        fu = cyc_pow(M, u)
        fu2 = cyc_pow(fu, u)
        fu3 = cyc_pow(fu2, u)
        t = frob(fu, 1); y3 = conj(t); free(t)
        t = frob(fu2, 1); t2 = mul(fu, t); y4 = conj(t2); free(t, t2)
        t = frob(fu3, 1); t2 = mul(fu3, t); y6 = conj(t2); free(t, t2, fu3, fu)
        y2 = frob(fu2, 2)
        y5 = conj(fu2); free(fu2)
        a1 = frob(M, 1); a2 = frob(M, 2); a3 = frob(M, 3)
        t = mul(a1, a2); y0 = mul(t, a3); free(t, a1, a2, a3)
        y1 = conj(M)
        # Devegili-Scott-Dahab chain
        t = csq(y6); t0 = mul(t, y4); free(t, y6, y4)
        t = mul(t0, y5); free(t0); t0 = t                  # t0 = y6^2 y4 y5
        t = mul(y3, y5); t1 = mul(t, t0); free(t, y3, y5)  # t1 = y3 y5 t0
        t = mul(t0, y2); free(t0, y2); t0 = t              # t0 = t0 y2
        t = csq(t1); free(t1); t1 = mul(t, t0); free(t)    # t1 = t1^2 t0
        t = csq(t1); free(t1); t1 = t                      # t1 = t1^2
        t = mul(t1, y1); free(t0, y1); t0 = t              # t0 = t1 y1
        t = mul(t1, y0); free(t1, y0); t1 = t              # t1 = t1 y0
        t = csq(t0); free(t0); t0 = t                      # t0 = t0^2
        gen.f12_mul(regs["FA"], t0, t1)
    else:
        x = cfg.u
        c = (x + 1) ** 2 // 3
        y0 = cyc_pow(M, c)
        t = cyc_pow(y0, x); a = conj(t); free(t)           # y0^x (x < 0)
        b = frob(y0, 1); y1 = mul(a, b); free(a, b, y0)    # y1 = y0^(x+p)
        t = cyc_pow(y1, x); t2 = cyc_pow(t, x); free(t)    # y1^(x^2)
        a = frob(y1, 2); b = mul(t2, a); free(t2, a)
        a = conj(y1); y2 = mul(b, a); free(a, b, y1)       # y1^(x^2+p^2-1)
        gen.f12_mul(regs["FA"], y2, M)
    # ---- out of Montgomery form: OUT = FA * RAW1 / R  (value < 2p, canonicalised by the kernel epilogue)
    OUT = gen.f12("OUT")
    gen.dot([(OUT[k][c], [(regs["FA"][k][c], gen.RAW1)]) for k in range(6) for c in range(2)])
    return regs["FA"], OUT


# ======================================================================================
# exact integer simulator (test infrastructure)
# ======================================================================================

class Sim:
    def __init__(self, gen: Gen):
        self.gen, self.cfg = gen, gen.cfg
        self.gs = [0] * gen._nslots
        self.maxv = 0

    def rd(self, s):
        return self.gs[s[1]] if s[0] == "g" else self.gen.cints[s[1]]

    def run(self, prog):
        cfg = self.cfg
        pinv = pow(cfg.p, -1, cfg.R)
        for pid in self.gen.programs[prog]:
            ph = self.gen.phases[pid]
            outs = []
            for dst, terms in ph.tasks:
                if ph.kind == "INV":
                    v = pow(self.rd(terms[0][0]) % cfg.p, -1, cfg.p)
                    outs.append((dst, v))
                    continue
                if ph.kind == "DOT":
                    t = sum((-1 if (len(tm_) > 2 and tm_[2] < 0) else 1) * self.rd(tm_[0]) * self.rd(tm_[1]) for tm_ in terms)
                    m = (-t * pinv) % cfg.R
                    v = (t + m * cfg.p) // cfg.R + getattr(ph, "K", 0) * cfg.p
                else:
                    v = 0
                    for s, coef, K in terms:
                        v += coef * self.rd(s) if coef > 0 else -coef * (K * cfg.p - self.rd(s))
                assert 0 <= v < cfg.R
                self.maxv = max(self.maxv, v)
                outs.append((dst, v))
            for dst, v in outs:
                self.gs[dst[1]] = v

    def set(self, slot, v):
        self.gs[slot[1]] = v

    def get(self, slot):
        return self.rd(slot)


def build_all(cfg):
    """Three independent slot files: "M" (16-lane MILLER, MUL_AB, MUL_BA), "F" (FINALEXP) and "P" (32-lane
    pipelined MILLER with signed terms: the latency path)."""
    gm, gf = Gen(cfg), Gen(cfg, lanes=32, tm=6, split=True, signed=True)
    gp = Gen(cfg, lanes=32, tm=6, signed=True)
    io = {}
    io["p_miller_in"], io["P_FA"] = build_miller_p(gp)
    io["P_GB"] = build_mulacc_p(gp, io["P_FA"])
    io["miller_in"], io["FA"] = build_miller(gm)
    _, io["FB"], io["GB"] = build_mulacc(gm)
    io["F_FA"], io["OUT"] = build_final_exp(gf)
    p = cfg.p
    # EXPORT (F file): OUT = FA out of Montgomery form, no exponentiation (raw Miller products for sharding)
    gf.begin("EXPORT")
    for k in range(6):
        for c in range(2):
            gf.set_ub(io["F_FA"][k][c], 64 * p)
    gf.dot([(io["OUT"][k][c], [(io["F_FA"][k][c], gf.RAW1)]) for k in range(6) for c in range(2)])
    # IMPORT_A / IMPORT_G (M file): plain-domain integers (wire bytes) -> Montgomery form in FA / GB
    io["RAWF"] = gm.f12("RAWF")
    for name, dst in (("IMPORT_A", io["FA"]), ("IMPORT_G", io["GB"])):
        gm.begin(name)
        for k in range(6):
            for c in range(2):
                gm.set_ub(io["RAWF"][k][c], p - 1)
        gm.dot([(dst[k][c], [(io["RAWF"][k][c], gm.R2)]) for k in range(6) for c in range(2)])
    return {"M": gm, "F": gf, "P": gp}, io


# ======================================================================================
# emit C tables
# ======================================================================================

def verify_all(cfg, gens, io):
    """Static (worst-case) bound verification of every program in execution order."""
    p = cfg.p
    gm, gf = gens["M"], gens["F"]
    raw, FA, FB, GB = io["miller_in"], io["FA"], io["FB"], io["GB"]
    rep = {}
    rep["MILLER"], ubm = gm.verify_program("MILLER", {s: p - 1 for s in raw.values()})
    f_ub = max(ubm[FA[k][c]] for k in range(6) for c in range(2))
    rep["P.MILLER"], ubp = gens["P"].verify_program("MILLER", {s: p - 1 for s in io["p_miller_in"].values()})
    assert max(ubp[io["P_FA"][k][c]] for k in range(6) for c in range(2)) <= TREE_P * p, "P Miller output exceeds the product-tree input bound"
    rep["P.MULACC"], ubx = gens["P"].verify_program("MULACC", {reg[k][c]: TREE_P * p for reg in (io["P_FA"], io["P_GB"]) for k in range(6) for c in range(2)})
    assert max(ubx[io["P_FA"][k][c]] for k in range(6) for c in range(2)) <= TREE_P * p, "product-tree invariant (P file)"
    # product trees: operands are Miller outputs or earlier products
    init = {reg[k][c]: 64 * p for reg in (FA, FB, GB) for k in range(6) for c in range(2)}
    assert f_ub <= 64 * p
    for nm in ("MUL_AB", "MUL_BA"):
        rep[nm], ubx = gm.verify_program(nm, init)
        out = FB if nm == "MUL_AB" else FA
        assert max(ubx[out[k][c]] for k in range(6) for c in range(2)) <= 64 * p
    FFA, OUT = io["F_FA"], io["OUT"]
    rep["FINALEXP"], ubf = gf.verify_program("FINALEXP", {FFA[k][c]: 64 * p for k in range(6) for c in range(2)})
    assert max(ubf[OUT[k][c]] for k in range(6) for c in range(2)) < 2 * p, "final output must be < 2p for the epilogue"
    rep["EXPORT"], ube = gf.verify_program("EXPORT", {FFA[k][c]: 64 * p for k in range(6) for c in range(2)})
    assert max(ube[OUT[k][c]] for k in range(6) for c in range(2)) < 2 * p
    for nm in ("IMPORT_A", "IMPORT_G"):
        rep[nm], _ = gm.verify_program(nm, {io["RAWF"][k][c]: p - 1 for k in range(6) for c in range(2)})
    return {k: v.bit_length() for k, v in rep.items()}


def emit_tables(path):
    out = ["// GENERATED by tools/gen_machine.py -- do not edit.", "#pragma once", "#include <cstdint>", '#include "arith.cuh"', "namespace bgls {", "namespace mtab {"]
    for cfg in (BN, BLS):
        gens, io = build_all(cfg)
        bits = verify_all(cfg, gens, io)
        print(cfg.name, "worst-case value bits per program:", bits, "of", cfg.W * cfg.L)
        for tag, gen in gens.items():
            n = "%s_%s" % (cfg.name, {"P": "MP"}.get(tag, tag))
            nsg = gen._nslots
            ref = lambda s, nsg=nsg: s[1] if s[0] == "g" else nsg + s[1]
            out.append("// ---- %s: W=%d L=%d, %d group slots, %d constant slots, %d phases" % (n, cfg.W, cfg.L, nsg, len(gen.cvals), len(gen.phases)))
            out.append("struct %s {" % n)
            out.append("  static constexpr int W = %d, L = %d, NSG = %d, NCONST = %d, NPHASE = %d;" % (cfg.W, cfg.L, nsg, len(gen.cvals), len(gen.phases)))
            out.append("  static constexpr int LANES = %d, TM = %d, REC = %d;  // lanes per group, max DOT terms, u16 per lane record" % (gen.lanes, gen.tm, 2 * gen.tm + 2))
            out.append("  static constexpr bool SIGNED = %s;  // DOT terms may be negative (bit 15 of the first operand), offset K p in the header" % ("true" if gen.signed else "false"))
            out.append("  static constexpr int NS = NSG + NCONST;  // slots of one group file: group slots, then a private copy of the constants")
            out.append("  static constexpr uint32_t N0 = 0x%xu;" % cfg.n0)
            out.append("  static constexpr int FP_BYTES = %d;" % (32 if cfg.name == "BN254" else 48))
            out.append("  HD static constexpr uint32_t p(int i) {")
            out.append("    constexpr uint32_t v[%d] = {%s};" % (cfg.L, ", ".join("0x%xu" % x for x in cfg.limbs(cfg.p))))
            out.append("    return v[i];")
            out.append("  }")
            for pname, prog in gen.programs.items():
                out.append("  static constexpr int %s_LEN = %d;" % (pname, len(prog)))
            if tag in ("M", "P"):
                raw = io["miller_in"] if tag == "M" else io["p_miller_in"]
                out.append("  static constexpr int IN_XP = %d, IN_YP = %d, IN_XQX = %d, IN_XQY = %d, IN_YQX = %d, IN_YQY = %d;" % tuple(
                    ref(raw[k]) for k in ("xP", "yP", "xQ.x", "xQ.y", "yQ.x", "yQ.y")))
                regs_ = (("FA", io["FA"]), ("FB", io["FB"]), ("GB", io["GB"]), ("RAWF", io["RAWF"])) if tag == "M" else (("FA", io["P_FA"]), ("GB", io["P_GB"]))
            else:
                regs_ = (("FA", io["F_FA"]), ("OUT", io["OUT"]))
            for nm, reg in regs_:
                flat = [ref(reg[k][c]) for k in range(6) for c in range(2)]
                assert flat == list(range(flat[0], flat[0] + 12)), (nm, flat)
                out.append("  static constexpr int %s0 = %d;  // 12 consecutive slots: (k, re/im), k = power of w" % (nm, flat[0]))
            out.append("  static constexpr int ONE = %d, ZERO = %d;" % (ref(gen.ONE), ref(gen.ZERO)))
            out.append("};")
            out.append("static const uint32_t %s_P[%d] = {%s};" % (n, cfg.L, ", ".join("0x%xu" % x for x in cfg.limbs(cfg.p))))
            out.append("static const uint32_t %s_CONST[%d][%d] = {" % (n, len(gen.cvals), cfg.L))
            for v in gen.cvals:
                out.append("  {%s}," % ", ".join("0x%xu" % x for x in v))
            out.append("};")
            # phase header: kind (0 DOT, 1 LIN) | T << 8 | ntasks << 16; record: REC = 2 TM + 2 u16 per lane:
            #   [0] dst (0xFFFF = idle), [1..TM] a / source slots, [1+TM..2TM] b slots / LIN coefficient words, [2TM+1] pad
            TM, REC = gen.tm, 2 * gen.tm + 2
            hdr, rec = [], []
            for ph in gen.phases:
                if getattr(ph, "triples", None) is not None:
                    # Karatsuba-lane DOT phase (kind 3): every term is (a1 + a2) * (b1 + b2); lanes 3j, 3j+1, 3j+2 hold
                    # the Q, P, S parts of triple j (P writes re, S writes im), then the plain tasks;
                    # record: [0] dst, [1 + 2t], [2 + 2t] = a1 (bit 15: minus), a2, [1 + TM + 2t], [2 + TM + 2t] = b1, b2
                    Z = ref(gen.ZERO)
                    ntr, half = len(ph.triples), TM // 2
                    lanes_ = []
                    for re, im, terms in ph.triples:
                        lanes_.append((0xFFFF, [(a[1], None, b[1], None, sg) for a, b, sg in terms]))            # Q
                        lanes_.append((ref(re), [(a[0], None, b[0], None, sg) for a, b, sg in terms]))            # P
                        lanes_.append((ref(im), [(a[0], a[1], b[0], b[1], sg) for a, b, sg in terms]))            # S
                    for dst, terms in ph.tasks[:ph.nplain]:
                        lanes_.append((ref(dst), [(t[0], None, t[1], None, (t[2] if len(t) > 2 else 1)) for t in terms]))
                    assert len(lanes_) <= gen.lanes
                    tk = max(len(terms) for _, terms in lanes_)     # terms actually present (<= half)
                    assert 0 < tk <= half
                    hdr.append(3 | (tk << 8) | (ntr << 16) | (getattr(ph, "K", 0) << 24))
                    for lane in range(gen.lanes):
                        r = [0xFFFF] + [Z] * (2 * TM) + [0]
                        if lane < len(lanes_):
                            r[0] = lanes_[lane][0]
                            for t, term in enumerate(lanes_[lane][1]):
                                a1, a2, b1, b2 = term[:4]
                                sgn = term[4] if len(term) > 4 else 1
                                r[1 + 2 * t] = ref(a1) | (0x8000 if sgn < 0 else 0)
                                r[2 + 2 * t] = ref(a2) if a2 is not None else Z
                                r[1 + TM + 2 * t] = ref(b1)
                                r[2 + TM + 2 * t] = ref(b2) if b2 is not None else Z
                        assert len(r) == REC
                        rec += r
                    continue
                hdr.append({"DOT": 0, "LIN": 1, "INV": 2}[ph.kind] | (ph.T << 8) | ((len(ph.tasks) & 0xFF) << 16) | (getattr(ph, "K", 0) << 24))
                for lane in range(gen.lanes):
                    r = [0xFFFF] + [ref(gen.ZERO)] * (2 * TM) + [0]
                    if ph.kind == "LIN":
                        for t in range(TM):
                            r[1 + TM + t] = 0
                        if gen.signed:
                            r[2 * TM + 1] = ref(gen.ZERO)   # idle lanes: a constant, never a slot another lane writes
                    if lane < len(ph.tasks):
                        dst, terms = ph.tasks[lane]
                        r[0] = ref(dst)
                        for t, term in enumerate(terms):
                            if ph.kind == "INV":
                                r[1] = ref(term[0])
                            elif ph.kind == "DOT":
                                r[1 + t], r[1 + TM + t] = ref(term[0]), ref(term[1])
                                assert r[1 + t] < 0x8000
                                if len(term) > 2 and term[2] < 0:
                                    r[1 + t] |= 0x8000   # minus sign: the interpreter negates the first operand
                            elif gen.signed:
                                slot, coef, K = term
                                r[1 + t] = ref(slot)
                                assert abs(coef) < 128
                                r[1 + TM + t] = coef & 0xFF      # two's complement coefficient
                            else:
                                slot, coef, K = term
                                r[1 + t] = ref(slot)
                                # |coef| in bits 0..6, sign in bit 7, KP constant index in bits 8..15
                                kp = gen.KP[K][1] if coef < 0 else 0
                                assert abs(coef) < 128 and kp < 256
                                r[1 + TM + t] = abs(coef) | (0x80 if coef < 0 else 0) | (kp << 8)
                        if ph.kind == "LIN" and gen.signed:
                            r[2 * TM + 1] = ref(ph.kt[lane])   # the task's offset constant (k * p)
                    assert len(r) == REC
                    rec += r
            out.append("static const uint32_t %s_PHASE_HDR[%d] = {%s};" % (n, len(hdr), ", ".join(str(h) for h in hdr)))
            out.append("static const uint16_t %s_PHASE_REC[%d] = {" % (n, len(rec)))
            for i in range(0, len(rec), REC):
                out.append("  " + ",".join(str(x) for x in rec[i:i + REC]) + ",")
            out.append("};")
            for pname, prog in gen.programs.items():
                out.append("static const uint16_t %s_PROG_%s[%d] = {%s};" % (n, pname, len(prog), ",".join(str(x) for x in prog)))
            out.append("struct %s_T {  // host-side accessors of the tables above" % n)
            out.append("  static const uint32_t* consts() { return &%s_CONST[0][0]; }" % n)
            out.append("  static const uint32_t* hdr() { return %s_PHASE_HDR; }" % n)
            out.append("  static const uint16_t* rec() { return %s_PHASE_REC; }" % n)
            for pname in gen.programs:
                out.append("  static const uint16_t* prog_%s() { return %s_PROG_%s; }" % (pname, n, pname))
            out.append("};")
    out += ["}  // namespace mtab", "}  // namespace bgls", ""]
    with open(path, "w") as f:
        f.write("\n".join(out))
    print("wrote", path)


if __name__ == "__main__":
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    emit_tables(os.path.join(root, "bgls_b200", "csrc", "machine_tables.cuh"))
