#!/usr/bin/env python
"""Condense `ncu -i rep --page raw --csv` dumps (gpurun_out/ncu_cfg*_raw.csv) into
the handful of counters DESIGN.md argues from; one block per captured launch.
    python profiles/summarize_ncu.py gpurun_out/ncu_cfg5a_raw.csv > profiles/r01_ncu_cfg5a.txt
"""
import csv
import sys

KEYS = [
    "gpu__time_duration.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__bytes_read.sum.per_second",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor_op_dmma.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_op_dmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__warps_eligible.avg.per_cycle_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic",
    "gpc__cycles_elapsed.avg.per_second", "dram__cycles_elapsed.avg.per_second",
]
STALL = "smsp__pcsamp_warps_issue_stalled_"


def main(path):
    rows = list(csv.reader(open(path)))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    names, units = rows[hdr], rows[hdr + 1]
    for r in rows[hdr + 2:]:
        if len(r) < len(names):
            continue
        d = dict(zip(names, r))
        u = dict(zip(names, units))
        print(f"kernel: {d['Kernel Name']}")
        for k in KEYS:
            if k in d:
                print(f"  {k:88s} {d[k]:>16s} {u[k]}")
        for k in names:
            if ("dmma" in k or "pipe_tensor" in k) and "pct" in k and k not in KEYS \
                    and d[k] not in ("0", "", "n/a"):
                print(f"  {k:88s} {d[k]:>16s} {u[k]}")
        stalls = sorted(((float(d[k] or 0), k[len(STALL):]) for k in names
                         if k.startswith(STALL) and not k.endswith("_not_issued")),
                        reverse=True)
        tot = sum(v for v, _ in stalls) or 1.0
        print("  issue-stall samples: " + ", ".join(f"{n} {100*v/tot:.0f}%" for v, n in stalls[:7]))
        print()


if __name__ == "__main__":
    main(sys.argv[1])
