#!/usr/bin/env python
"""profiles/<name>.csv from an .ncu-rep (ncu --set full, one kernel): the launch, time, DRAM traffic,
pipe utilisation, issue statistics and stall reasons per issued instruction.
  python scripts/summarize_ncu.py gpurun_out/prof.ncu-rep profiles/r01b_ncu_f32.csv"""
import csv
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__cycles_active.avg",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    rows = [r for r in rows if r and not r[0].startswith("==")]
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit", "value"])
        w.writerow(["Kernel Name", "", d["Kernel Name"][1]])
        for k in KEEP:
            if k in d:
                w.writerow([k, d[k][0], d[k][1]])
        for h, (u, v) in sorted(d.items()):
            if "issue_stalled" in h and h.endswith("per_issue_active.ratio"):
                w.writerow([h, u, v])
    print("wrote", out)


if __name__ == "__main__":
    main()
