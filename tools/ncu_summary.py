#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed): key raw metrics per captured kernel + top source hot spots.
usage: python tools/ncu_summary.py gpurun_out/prof_X.ncu-rep [kernel-regex] > profiles/rN_X.txt"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sectors.sum", "lts__t_sectors.sum.per_second",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_hit_rate.pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_active.avg.per_cycle_active",
        "smsp__warps_eligible.avg.per_cycle_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma_type_fp16.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio"]


def ncu(*args):
    return subprocess.run(["ncu", *args], capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    kre = sys.argv[2] if len(sys.argv) > 2 else None
    raw = list(csv.reader(io.StringIO(ncu("-i", rep, "--page", "raw", "--csv"))))
    hdr, units = raw[0], raw[1]
    print("# %s  (ncu --set full --clock-control none; per-launch values)" % rep)
    for r in raw[2:]:
        print("\n== kernel: %s   grid %s block %s" % (r[hdr.index("Kernel Name")], r[hdr.index("Grid Size")] if "Grid Size" in hdr else "?",
                                                      r[hdr.index("Block Size")] if "Block Size" in hdr else "?"))
        for k in KEYS:
            if k in hdr:
                print("%-86s %s %s" % (k, r[hdr.index(k)], units[hdr.index(k)]))
    args = ["-i", rep, "--page", "source", "--csv"]
    if kre:
        args += ["--kernel-name", "regex:" + kre]
    src = list(csv.reader(io.StringIO(ncu(*args))))
    if len(src) > 2:
        h = src[1]
        ie, isamp, isrc = h.index("Instructions Executed"), h.index("# Samples"), h.index("Source")
        data = [(int(r[ie]), int(r[isamp]), r[isrc].strip()) for r in src[2:] if len(r) > ie and r[ie].isdigit()]
        ti, ts = sum(d[0] for d in data), sum(d[1] for d in data)
        print("\n== source page (all captured launches): %d SASS lines, %d instructions executed, %d samples" % (len(data), ti, ts))
        print("   64-instruction windows holding >= 1.5%% of the executed instructions:")
        for k in range(0, len(data), 64):
            seg = data[k:k + 64]
            s, m = sum(d[0] for d in seg), sum(d[1] for d in seg)
            if s >= 0.015 * ti:
                print("   [%5d..%5d] inst %5.1f%%  samples %5.1f%%   %s" % (k, k + len(seg) - 1, 100.0 * s / ti, 100.0 * m / max(ts, 1), seg[0][2][:60]))


if __name__ == "__main__":
    main()
