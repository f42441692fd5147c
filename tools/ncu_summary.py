#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed) into the few numbers DESIGN.md / bench.py cite.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/sim_r01.json
"""
import csv
import io
import json
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.avg",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
]
STALLS = "smsp__average_warps_issue_stalled_"


def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    launches = []
    for r in data:
        d = {"kernel": r[hdr.index("Kernel Name")]}
        for i, h in enumerate(hdr):
            if h in KEYS or (h.startswith(STALLS) and h.endswith("_per_issue_active.ratio")):
                try:
                    v = float(r[i])
                except ValueError:
                    continue
                if v != v:
                    continue
                key = h.replace(STALLS, "stall_").replace("_per_issue_active.ratio", "")
                d[key] = {"value": v, "unit": units[i]}
        launches.append(d)
    json.dump({"report": rep, "launches": launches}, open(out, "w"), indent=1)
    for d in launches:
        print(d["kernel"])
        for k, v in d.items():
            if k != "kernel":
                print(f"  {k:75s} {v['value']:.6g} {v['unit']}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
