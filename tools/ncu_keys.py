"""Print the handful of ncu metrics the design discussion uses from a --set full .ncu-rep (raw page)."""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fmaheavy", "sm__inst_executed_pipe_fmalite", "sm__inst_executed_pipe_fma.",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "smsp__average_warps_issue_stalled", "sm__cycles_active.avg",
        "smsp__cycles_active.avg", "sm__inst_executed_pipe_lsu", "sm__inst_executed.avg.per_cycle_active",
        "smsp__inst_executed_pipe", "sm__inst_executed_pipe_imad", "sm__pipe_imad"]


def main(path, idx=0):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2 + idx]
    for h, u, v in zip(hdr, units, vals):
        if any(k in h for k in KEYS):
            try:
                if float(v) == 0:
                    continue
            except ValueError:
                pass
            print(f"{h:95s} {u:16s} {v}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0)
