#!/usr/bin/env python
"""profiles/r02_traffic.json from an `ncu --set full` report: per-launch DRAM bytes and issue statistics of the kernels
bench.py's roofline object quotes (bench.py cannot measure them itself: no profiler inside a timed run).

    python tools/ncu_traffic.py gpurun_out/<report>.ncu-rep --workload c2_16m --command "<the profiled command>" [--merge]

Kernel name -> bench stage: fgpu_k_<f> -> func:<f>; the build kernels are recorded under their own names.  The commit
the report was taken at and the profiled command are stored beside the numbers."""
import argparse, csv, io, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = {"dram__bytes_read.sum": "dram_read_bytes", "dram__bytes_write.sum": "dram_write_bytes", "gpu__time_duration.sum": "ncu_duration_us",
        "smsp__inst_executed.sum": "warp_instructions", "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_frac",
        "smsp__thread_inst_executed_per_inst_executed.ratio": "active_lanes_per_instruction",
        "l1tex__t_sector_hit_rate.pct": "l1_sector_hit_rate", "lts__t_sector_hit_rate.pct": "l2_sector_hit_rate",
        "launch__registers_per_thread": "registers", "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_frac",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed": "l1_data_pipe_frac",
        "l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed": "lsu_writeback_frac",
        "SM_A.TriageCompute.l1tex__data_pipe_lsu_wavefronts.avg": "l1_wavefronts_per_sm",
        "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum": "global_load_requests"}
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "us": 1.0, "ms": 1e3, "ns": 1e-3, "s": 1e6, "%": 0.01}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("--workload", required=True)
    ap.add_argument("--command", default="")
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r02_traffic.json"))
    ap.add_argument("--merge", action="store_true")
    a = ap.parse_args()
    raw = subprocess.run(["ncu", "-i", a.report, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    kn = hdr.index("Kernel Name")
    commit = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], stdout=subprocess.PIPE, text=True).stdout.strip()
    out = json.load(open(a.out)) if a.merge and os.path.exists(a.out) else {}
    out["_comment"] = ("per-launch figures from `ncu --set full --clock-control none` reports (tools/ncu_traffic.py); bench.py copies "
                       "them into roofline.traffic / roofline.issue -- they are not measured inside bench.py")
    wl = out.setdefault(a.workload, {})
    seen = set()
    for r in rows[2:]:
        name = r[kn].split("(")[0].replace("void ", "").replace("fgpu::", "")
        stage = "func:" + name[len("fgpu_k_"):] if name.startswith("fgpu_k_") else name.split("<")[0]
        rec = {"kernel": r[kn][:80], "commit": commit, "command": a.command, "report": os.path.basename(a.report)}
        for k, dst in KEYS.items():
            if k in hdr:
                i = hdr.index(k)
                try:
                    rec[dst] = float(r[i].replace(",", "")) * SCALE.get(units[i], 1.0)
                except ValueError:
                    pass
        for k in ("dram_read_bytes", "dram_write_bytes", "warp_instructions", "registers"):
            if k in rec:
                rec[k] = int(round(rec[k]))
        if stage not in seen:                # the first launch of each kernel in the report (replaces an older capture)
            seen.add(stage)
            wl[stage] = rec
    json.dump(out, open(a.out, "w"), indent=1)
    print("wrote", a.out, {k: sorted(v) for k, v in out.items() if not k.startswith("_")})


if __name__ == "__main__":
    main()
