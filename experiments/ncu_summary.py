"""Summarise `ncu --set full` reports (read on the CPU box with `ncu -i ... --page raw --csv`)."""
import csv, subprocess, sys, io, json

KEYS = [
 "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
 "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
 "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
 "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
 "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
 "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
 "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
 "smsp__inst_executed.sum", "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_xu.sum", "sm__inst_executed_pipe_alu.sum",
 "sm__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_fp64.sum",
 "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum",
 "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ldgsts.sum",
 "smsp__average_warp_latency_issue_stalled_math_pipe_throttle.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
 "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
 "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
 "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
 "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
 "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
 "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
 "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
 "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
 "local_load", "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum", "sm__cycles_elapsed.avg", "smsp__cycles_active.avg",
 "sm__cycles_active.avg", "gpc__cycles_elapsed.avg.per_second", "sm__cycles_elapsed.avg.per_second",
]

def load(path):
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    return [dict((h, (u, v)) for h, u, v in zip(hdr, units, r)) for r in rows[2:]]

if __name__ == "__main__":
    for path in sys.argv[1:]:
        for k in load(path):
            print("## %s  kernel=%s grid=%s block=%s" % (path, k.get("Kernel Name", ("", ""))[1], k.get("Grid Size", ("", ""))[1], k.get("Block Size", ("", ""))[1]))
            for key in KEYS:
                if key in k:
                    print("%-90s %-14s %s" % (key, k[key][0], k[key][1]))
            print()
