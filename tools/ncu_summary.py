"""Prints the handful of ncu metrics that matter for the integer-pipe kernels from a .ncu-rep
(ncu -i <rep> --page raw --csv).   python tools/ncu_summary.py gpurun_out/prof.ncu-rep [kernel-substring]"""
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma", "sm__inst_executed_pipe_alu",
    "sm__inst_executed_pipe_lsu", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fmaheavy_cycles_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__average_warps_issue_stalled", "smsp__average_warp_latency_issue_stalled", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__cycles_elapsed.avg",
    "smsp__cycles_active.avg", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__sass_inst_executed_op_shared", "smsp__inst_executed_op_shared",
]


def main():
    rep = sys.argv[1]
    sub = sys.argv[2] if len(sys.argv) > 2 else ""
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[0]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else ""
        if sub not in name:
            continue
        print("==", name[:100])
        for h, v in zip(hdr, r):
            if any(h.startswith(k) for k in KEYS) and v not in ("", "0", "n/a"):
                print("  %-90s %s" % (h, v))


if __name__ == "__main__":
    main()
