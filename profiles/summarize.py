"""Turn gpurun_out/*.ncu-rep (+ the launch list) into the committed summaries under profiles/.
Run here (no GPU needed): python profiles/summarize.py"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
SRC = os.path.join(ROOT, "gpurun_out")
KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum", "sm__cycles_elapsed.avg.per_second", "smsp__inst_executed.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "lts__t_sectors_srcunit_tex_op_read.sum",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum",
    "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum.per_second",
]


def rows_of(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = {"kernel": r[hdr.index("Kernel Name")]}
        for k in KEYS:
            if k in hdr:
                d[k] = f"{r[hdr.index(k)]} {units[hdr.index(k)]}".strip()
        out.append(d)
    return out


def main():
    for rnd in ("r1", "r2"):
        summary = {}
        for name in sorted(os.listdir(SRC)):
            if name.startswith("prof_" + rnd) and name.endswith(".ncu-rep"):
                summary[name[:-8]] = rows_of(os.path.join(SRC, name))
        if summary:
            with open(os.path.join(OUT, rnd + "_ncu_full_summary.json"), "w") as f:
                json.dump(summary, f, indent=1)
    # launch list: per-kernel totals and shares of one timed step
    ll = os.path.join(SRC, "launches_r1.csv")
    if os.path.exists(ll):
        rows = [r for r in csv.reader(open(ll)) if len(r) > 5]
        start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
        rows = rows[start:]
        hdr = rows[0]
        ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
        launches = [(r[ki], float(r[vi].replace(",", ""))) for r in rows[1:] if r[hdr.index("Metric Name")] == "gpu__time_duration.sum"]
        with open(os.path.join(OUT, "r1_launch_list.csv"), "w") as f:
            f.write("index,kernel,duration_ns\n")
            for i, (k, v) in enumerate(launches):
                short = k.split("(")[0].replace(",", ";")
                f.write(f"{i},{short},{v:.0f}\n")
        print("launches:", len(launches))
    print(json.dumps({k: [r["kernel"][:30] + " " + r.get("gpu__time_duration.sum", "") for r in v] for k, v in summary.items()}, indent=1))


if __name__ == "__main__":
    main()
