#!/usr/bin/env python3
"""Selected metrics of every kernel instance in an `ncu --set full` report, as metric,unit,value rows:
    python tools/ncu_full_summary.py REPORT.ncu-rep [instance index] > profiles/rN_ncu_full_<workload>_<kernel>.csv
(needs the `ncu` binary; reads the report through `ncu -i ... --page raw --csv`)."""
import csv
import io
import subprocess
import sys

WANT = ["Kernel Name", "Block Size", "Grid Size",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes.sum.per_second", "gpu__time_duration.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "derived__smsp__inst_executed_op_local_st_pct", "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum"]


def main():
    rep = sys.argv[1]
    which = int(sys.argv[2]) if len(sys.argv) > 2 else None
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    out = csv.writer(sys.stdout)
    out.writerow(["metric", "unit", "value"])
    for k, r in enumerate(rows[2:]):
        if which is not None and k != which:
            continue
        for m in WANT:
            if m in idx:
                out.writerow([m, units[idx[m]], r[idx[m]]])
        out.writerow([])


if __name__ == "__main__":
    main()
