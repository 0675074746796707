#!/usr/bin/env python
"""Summarise an ``ncu --set full`` capture into the JSON kept under ``profiles/``.

    python tools/ncu_summary.py gpurun_out/prof_x.ncu-rep profiles/r02_x_summary.json [launch index]

Reads the report here (no GPU needed) with ``ncu -i ... --page raw --csv`` and keeps the metrics
DESIGN.md and bench.py's ``roofline.traffic`` cite: duration, DRAM bytes, L2 hit rate, tensor-pipe and
issue utilisation, shared-memory wavefronts, launch geometry.
"""
import csv
import io
import json
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__m_xbar2l1tex_read_bytes.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_tensor.sum", "smsp__issue_active.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum", "smsp__inst_executed_op_shared_atom.sum",
    "launch__registers_per_thread", "launch__block_size", "launch__grid_size", "launch__cluster_size",
    "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.max", "smsp__cycles_active.avg",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    idx = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True,
                         text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    header, units = rows[0], rows[1]
    data = rows[2 + idx]
    res = {"_report": rep, "_launch_index": idx, "_launches_in_report": len(rows) - 2}
    for name, unit, val in zip(header, units, data):
        if name == "Kernel Name" or name in KEEP:
            res[name] = {"unit": unit, "value": val}
    with open(out, "w") as f:
        json.dump(res, f, indent=1)
    dur = res.get("gpu__time_duration.sum", {})
    print(out, res.get("Kernel Name", {}).get("value", "?")[:80], dur.get("value"), dur.get("unit"))


if __name__ == "__main__":
    main()
