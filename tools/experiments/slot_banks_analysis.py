"""Experiment: shared-memory wavefronts per warp-wide slot access of every generated program under the current slot
numbering (1.0 = conflict free); see slot_banks_anneal.py.      python tools/experiments/slot_banks_analysis.py"""
import sys, collections, time
sys.path.insert(0, __import__('os').path.join(__import__('os').path.dirname(__import__('os').path.abspath(__file__)), '..'))
import gen_machine as G

def instrs_of_phase(gen, ph):
    """list of (kind, [slot refs per lane]) : one entry per warp-wide LDS/STS group (each is L instructions)"""
    Z = gen.ZERO
    TM = gen.tm
    out = []
    lanes = gen.lanes
    if getattr(ph, "triples", None) is not None:
        lanes_ = []
        for re, im, terms in ph.triples:
            lanes_.append((None, [(a[1], None, b[1], None) for a, b, sg in terms]))
            lanes_.append((re, [(a[0], None, b[0], None) for a, b, sg in terms]))
            lanes_.append((im, [(a[0], a[1], b[0], b[1]) for a, b, sg in terms]))
        for dst, terms in ph.tasks[:ph.nplain]:
            lanes_.append((dst, [(t[0], None, t[1], None) for t in terms]))
        tk = max(len(t) for _, t in lanes_)
        for t in range(tk):
            for f in range(4):
                v = []
                for ln in range(lanes):
                    if ln < len(lanes_) and t < len(lanes_[ln][1]) and lanes_[ln][1][t][f] is not None:
                        v.append(lanes_[ln][1][t][f])
                    else:
                        v.append(Z)
                out.append(("ld", v))
        out.append(("st", [lanes_[ln][0] if ln < len(lanes_) else None for ln in range(lanes)]))
        return out
    if ph.kind == "DOT":
        T = ph.T
        for t in range(T):
            for f in range(2):
                v = []
                for ln in range(lanes):
                    if ln < len(ph.tasks) and t < len(ph.tasks[ln][1]):
                        v.append(ph.tasks[ln][1][t][f])
                    else:
                        v.append(Z)
                out.append(("ld", v))
        out.append(("st", [ph.tasks[ln][0] if ln < len(ph.tasks) else None for ln in range(lanes)]))
    elif ph.kind == "LIN":
        T = ph.T
        if gen.signed:
            out.append(("ld", [ph.kt[ln] if ln < len(ph.tasks) else Z for ln in range(lanes)]))
        for t in range(T):
            v = []
            for ln in range(lanes):
                if ln < len(ph.tasks) and t < len(ph.tasks[ln][1]):
                    v.append(ph.tasks[ln][1][t][0])
                else:
                    v.append(Z)
            out.append(("ld", v))
        out.append(("st", [ph.tasks[ln][0] if ln < len(ph.tasks) else None for ln in range(lanes)]))
    return out

def cost(instrs_w, pos):
    tot = 0; ideal = 0
    for w, v in instrs_w:
        banks = collections.defaultdict(set)
        for s in v:
            if s is None: continue
            banks[pos(s) & 31].add(s)
        tot += w * max(len(b) for b in banks.values())
        ideal += w
    return tot, ideal

if __name__ == "__main__":
    for cfg in (G.BN, G.BLS):
        gens, io = G.build_all(cfg)
        for tag, gen in gens.items():
            nsg = gen._nslots
            pos = lambda s: s[1] if s[0] == "g" else nsg + s[1]
            for pname, prog in gen.programs.items():
                cnt = collections.Counter(prog)
                iw = []
                for pid, w in cnt.items():
                    for kind, v in instrs_of_phase(gen, gen.phases[pid]):
                        iw.append((w, v))
                c, i = cost(iw, pos)
                print(cfg.name, tag, pname, "phases", len(prog), "wavefront factor %.3f" % (c / i))
