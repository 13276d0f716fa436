"""Experiment (not part of the product, result in DESIGN.md section 5): renumber the slots of each slot file by seeded
annealing so that the lanes of a warp-wide operand fetch hit different shared-memory banks.  Removes the bank conflicts
(1.49 -> 1.02 wavefronts per access with all slots free) but does not change any kernel time on B200, so the generator
keeps its natural numbering.      python tools/experiments/slot_banks_anneal.py"""
import collections
import os
import random
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import gen_machine as G
from slot_banks_analysis import instrs_of_phase

def optimize(gen, fixed, weights=None, iters=60000, seed=1):
    nsg = gen._nslots
    pos0 = lambda s: s[1] if s[0] == "g" else nsg + s[1]
    merged = collections.Counter()
    for pname, prog in gen.programs.items():
        wprog = (weights or {}).get(pname, 1)
        for pid, w in collections.Counter(prog).items():
            for kind, v in instrs_of_phase(gen, gen.phases[pid]):
                key = tuple(sorted({pos0(s) for s in v if s is not None}))
                merged[key] += w * wprog
    instrs = [(list(k), w) for k, w in merged.items()]
    by_slot = collections.defaultdict(list)
    for idx, (k, w) in enumerate(instrs):
        for s in k:
            by_slot[s].append(idx)
    perm = list(range(nsg + len(gen.cvals)))   # slot -> position
    def icost(idx):
        k, w = instrs[idx]
        cnt = [0] * 32
        for s in k:
            cnt[perm[s] & 31] += 1
        return w * max(cnt)
    costs = [icost(i) for i in range(len(instrs))]
    total = sum(costs); ideal = sum(w for _, w in instrs)
    free = [s for s in range(nsg) if s not in fixed]
    rng = random.Random(seed)
    start = total
    T0 = total / len(instrs) * 0.5
    for it in range(iters):
        T = T0 * (1 - it / iters) ** 2 + 1e-9
        x, y = rng.sample(free, 2)
        if (perm[x] & 31) == (perm[y] & 31):
            continue
        aff = set(by_slot[x]) | set(by_slot[y])
        old = sum(costs[i] for i in aff)
        perm[x], perm[y] = perm[y], perm[x]
        newc = {i: icost(i) for i in aff}
        new = sum(newc.values())
        d = new - old
        if d <= 0 or rng.random() < pow(2.718281828, -d / T):
            for i, c in newc.items():
                costs[i] = c
            total += d
        else:
            perm[x], perm[y] = perm[y], perm[x]
    return perm, start / ideal, total / ideal

if __name__ == "__main__":
    for cfg in (G.BN, G.BLS):
        gens, io = G.build_all(cfg)
        for tag, gen in gens.items():
            if tag == "P":
                blocks = [io["P_FA"], io["P_GB"]]
            elif tag == "M":
                blocks = [io["FA"], io["FB"], io["GB"], io["RAWF"]]
            else:
                blocks = [io["F_FA"], io["OUT"]]
            fixed = {reg[k][c][1] for reg in blocks for k in range(6) for c in range(2)}
            t = time.time()
            perm, a, b = optimize(gen, fixed)
            print(cfg.name, tag, "factor %.3f -> %.3f" % (a, b), "%.1fs" % (time.time() - t))
