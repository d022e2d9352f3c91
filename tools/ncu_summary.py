"""Print the key ncu metrics of a .ncu-rep (developer tool; reads with `ncu -i ... --page raw --csv`)."""
import csv
import subprocess
import sys

KEYS = """gpu__time_duration.sum
smsp__inst_executed.sum
smsp__issue_active.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active
sm__warps_active.avg.pct_of_peak_sustained_active
l1tex__data_pipe_lsu_wavefronts_mem_shared.sum
l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum
l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed
dram__bytes_read.sum
dram__bytes_write.sum
launch__registers_per_thread
launch__occupancy_limit_shared_mem
launch__occupancy_limit_registers
launch__occupancy_limit_warps
launch__grid_size
sm__throughput.avg.pct_of_peak_sustained_elapsed
smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio
smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio
smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio
smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio
smsp__average_warps_issue_stalled_wait_per_issue_active.ratio
smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio
smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio
smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio
smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio
smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio
smsp__average_warps_issue_stalled_selected_per_issue_active.ratio""".split()


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, u = rows[0], rows[1]
    for v in rows[2:]:
        print("kernel:", v[h.index("Kernel Name")][:60])
        for i, n in enumerate(h):
            if n in KEYS:
                print(f"  {n:88s} {v[i]:>16s} {u[i]}")


if __name__ == "__main__":
    main(sys.argv[1])
