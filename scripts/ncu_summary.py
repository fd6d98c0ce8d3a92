#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu --set full) into a small text table: one block per kernel launch with the
counters DESIGN.md / bench.py quote (duration, DRAM bytes, pipe utilisation, issue slots, stall reasons,
occupancy, registers).  usage: python scripts/ncu_summary.py file.ncu-rep > profiles/xxx.txt"""
import csv
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("lts__t_bytes.sum", "L2 bytes"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__waves_per_multiprocessor", "waves/SM"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__inst_executed_pipe_fmaheavy.sum", "FMA-heavy pipe instructions"),
    ("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active", "FMA-heavy pipe active %"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe active %"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "FP64 pipe active %"),
    ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "ALU pipe active %"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU (MUFU) pipe %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
]
STALL = "smsp__average_warps_issue_stalled_%s_per_issue_active.ratio"
STALLS = ["long_scoreboard", "short_scoreboard", "wait", "barrier", "membar", "math_pipe_throttle", "mio_throttle",
          "lg_throttle", "branch_resolving", "dispatch_stall", "no_instruction", "not_selected", "sleeping"]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(l for l in out.splitlines() if l.startswith('"')))
    hdr, units = rows[0], rows[1]
    print("# %s  (ncu --set full --clock-control none; per launch)" % rep.split("/")[-1])
    for r in rows[2:]:
        g = lambda k: (r[hdr.index(k)], units[hdr.index(k)]) if k in hdr else None
        print("\n== %s" % g("Kernel Name")[0][:110])
        for k, label in WANT:
            v = g(k)
            if v and v[0] != "":
                print("  %-32s %s %s" % (label, v[0], v[1]))
        st = [(s, g(STALL % s)) for s in STALLS]
        print("  stalled warps per issue:        " + ", ".join("%s %.2f" % (s, float(v[0])) for s, v in st if v and v[0] not in ("", "n/a") and float(v[0]) >= 0.05))


if __name__ == "__main__":
    main()
