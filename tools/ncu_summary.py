#!/usr/bin/env python
"""One-screen summary of an `ncu --set full` report: duration, instruction and memory-pipe utilisation, stall reasons.

    python tools/ncu_summary.py gpurun_out/<report>.ncu-rep [more reports ...]
"""
import csv, io, subprocess, sys

WANT = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__warps_eligible.avg.per_cycle_active"]

for path in sys.argv[1:]:
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    h, u = rows[0], rows[1]
    for r in rows[2:]:
        print("==", path, "|", r[h.index("Kernel Name")][:70])
        for k in WANT:
            if k in h:
                print("  %-78s %s %s" % (k, r[h.index(k)], u[h.index(k)]))
        stalls = {}
        for i, k in enumerate(h):
            if k.startswith("smsp__pcsamp_warps_issue_stalled_") and not k.endswith("_not_issued"):
                try:
                    stalls[k[len("smsp__pcsamp_warps_issue_stalled_"):]] = float(r[i])
                except ValueError:
                    pass
        tot = sum(stalls.values()) or 1.0
        print("  stall samples:", ", ".join("%s %.0f%%" % (k, 100 * v / tot) for k, v in sorted(stalls.items(), key=lambda t: -t[1])[:6]))
