"""Text summary of an .ncu-rep (one kernel): the raw-page metrics the DESIGN.md numbers come from and the executed
instruction mix by opcode from the source page.      python tools/ncu_summary.py <file.ncu-rep> [units per launch] > summary.txt"""
import collections
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_warps", "smsp__inst_executed.sum",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second"]


def page(rep, name):
    return list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout.splitlines()))


def main():
    rep = sys.argv[1]
    units = float(sys.argv[2]) if len(sys.argv) > 2 else None
    rows = page(rep, "raw")
    hdr, unit, vals = rows[0], rows[1], rows[2]
    m = dict(zip(hdr, vals))
    u = dict(zip(hdr, unit))
    print("== %s" % m.get("Kernel Name", "?")[:160])
    for w in WANT:
        if w in m:
            print("  %-86s %s %s" % (w, m[w], u.get(w, "")))
    for k in sorted(m):
        if "issue_stalled" in k and k.endswith("_per_issue_active.ratio") and "not_issued" not in k:
            print("  %-86s %s" % (k, m[k]))
    src = page(rep, "source")
    if len(src) > 2:
        h = src[1]
        ix = {x: i for i, x in enumerate(h)}
        mix = collections.Counter()
        for r in src[2:]:
            if len(r) != len(h):
                continue
            s = r[ix["Source"]].split()
            if not s:
                continue
            op = s[1] if s[0].startswith("@") and len(s) > 1 else s[0]
            op = op.rstrip(";")
            key = "IMAD.WIDE" if op.startswith("IMAD.WIDE") else op.split(".")[0] if not op.startswith("IMAD") else op.split(".U32")[0]
            mix[key] += int(r[ix["Instructions Executed"]] or 0)
        tot = sum(mix.values())
        print("  -- executed warp instructions by opcode (source page), total %d" % tot)
        for k, v in mix.most_common(14):
            print("     %-14s %12d  %.3f" % (k, v, v / tot) + ("   %.0f per unit" % (v / units) if units else ""))
        if units:
            print("     warp instructions per unit: %.0f" % (tot / units))


if __name__ == "__main__":
    main()
