"""Summary of an `ncu --set full` capture of one dense MSM accumulation (k_aff_forward / k_aff_invert / k_aff_backward x
rounds + k_msm_accum): per-kernel time, DRAM bytes and pipe / stall figures as text, and the per-accumulation DRAM
traffic (dram__bytes_read.sum + dram__bytes_write.sum over all of its kernels) as JSON -- bench.py's `roofline.traffic`.
  python tools/ncu_accum_summary.py <report.ncu-rep> <out.json>"""
import csv
import json
import subprocess
import sys


def num(v):
    try:
        return float(v.replace(",", ""))
    except ValueError:
        return None


def main(rep, out_json):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}

    def get(r, name, scale_to=None):
        v = num(r[col[name]]) if name in col else None
        if v is None:
            return None
        u = units[col[name]].lower()
        if scale_to == "bytes":
            v *= {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
        if scale_to == "ms":
            v *= {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1, "msecond": 1, "second": 1e3, "s": 1e3}.get(u, 1)
        return v

    keys = ["launch__registers_per_thread", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "sm__inst_executed.avg.per_cycle_active",
            "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio"]
    total_ms = total_bytes = 0.0
    kernels = []
    for r in rows[2:]:
        name = r[col["Kernel Name"]].split("(")[0]
        ms = get(r, "gpu__time_duration.sum", "ms") or 0.0
        rd = get(r, "dram__bytes_read.sum", "bytes") or 0.0
        wr = get(r, "dram__bytes_write.sum", "bytes") or 0.0
        total_ms += ms
        total_bytes += rd + wr
        kernels.append({"kernel": name, "ms": ms, "dram_read_bytes": rd, "dram_write_bytes": wr})
        print(f"--- {name}")
        print(f"  gpu__time_duration {ms:.3f} ms   dram read {rd / 1e9:.3f} GB  write {wr / 1e9:.3f} GB  "
              f"({(rd + wr) / 1e9 / (ms / 1e3) / 1e3 if ms else 0:.2f} TB/s)")
        for k in keys:
            if k in col:
                print(f"  {k:90s} {r[col[k]]}")
    print(f"=== one accumulation: {total_ms:.3f} ms summed over {len(kernels)} launches (cold cache, serialised), "
          f"{total_bytes / 1e9:.2f} GB of DRAM traffic")
    with open(out_json, "w") as f:
        json.dump({"what": "ncu --set full, dense 3-commitment MSM batch at 2^21 (tools/msm_affine_probe.py 21 3 1 3), all kernels of "
                           "the accumulation step: 3 x (k_aff_forward, k_aff_invert, k_aff_backward) + k_msm_accum",
                   "dram_bytes_per_launch": total_bytes, "ms_sum_under_ncu": total_ms, "kernels": kernels}, f, indent=1)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
