#!/usr/bin/env python
"""Turn one `ncu --set full` capture (.ncu-rep) into the short text summary kept in this directory.

usage: python profiles/summarize_ncu.py gpurun_out/r1_prof_fm2b.ncu-rep > profiles/r1_ncu_fm2b.txt
Reads the report through `ncu -i ... --page raw --csv` (no GPU needed).
"""
import csv
import io
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__inst_executed.avg.per_cycle_elapsed", "sm__inst_executed.avg.per_cycle_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
]
STALL = "smsp__average_warps_issue_stalled_"


def main(path):
    if path.endswith(".csv"):                 # the raw page already exported on the GPU box (ncu -i ... --page raw --csv)
        rows = list(csv.reader(open(path)))
    else:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True)
        rows = list(csv.reader(io.StringIO(out.stdout)))
    names, units = rows[0], rows[1]
    for vals in rows[2:]:
        rec = dict(zip(names, zip(units, vals)))
        print("kernel:", rec["Kernel Name"][1])
        for k in KEEP:
            if k in rec:
                print(f"  {k} [{rec[k][0]}] = {rec[k][1]}")
        stalls = {k[len(STALL):-len("_per_issue_active.ratio")]: float(v[1] or 0) for k, v in rec.items()
                  if k.startswith(STALL) and k.endswith("_per_issue_active.ratio")}
        tot = sum(stalls.values())
        print("  warp states per issue slot (share of resident-warp time, %):")
        for k, v in sorted(stalls.items(), key=lambda kv: -kv[1]):
            if v > 0.004:
                print(f"    {k:22s} {v:7.3f}  {100 * v / tot:5.1f}%")


if __name__ == "__main__":
    main(sys.argv[1])
