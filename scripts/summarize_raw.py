"""Turn an `ncu -i X.ncu-rep --page raw --csv` dump with SEVERAL captured launches into one small
markdown table (one column group per launch) under profiles/.  Usage:
    python scripts/summarize_raw.py <raw.csv> <out.md> <title> [note]"""
import csv
import sys

KEYS = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio"]
SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}


def main():
    raw, out, title = sys.argv[1], sys.argv[2], sys.argv[3]
    note = sys.argv[4] if len(sys.argv) > 4 else ""
    rows = list(csv.reader(open(raw)))
    names, units, data = rows[0], rows[1], rows[2:]
    ci = {n: i for i, n in enumerate(names)}
    with open(out, "w") as f:
        f.write("# %s\n\n%s\n\nDurations under the profiler (cold caches, serialised, ~40 replays) are not bench "
                "values.  Bytes in MB, durations in ms.\n\n" % (title, note))
        for r in data:
            if len(r) < len(names):
                continue
            f.write("## `%s`\n\n| metric | value |\n|---|---:|\n" % r[ci["Kernel Name"]][:110])
            dr = dw = dur = None
            for k in KEYS:
                if k not in ci:
                    continue
                v, u = r[ci[k]], units[ci[k]]
                try:
                    x = float(v.replace(",", ""))
                except ValueError:
                    continue
                if u in ("byte", "Kbyte", "Mbyte", "Gbyte"):
                    x = x * SCALE[u] / 1e6
                    u = "MB"
                if k == "gpu__time_duration.sum":
                    x = x * SCALE.get(u, 1.0)
                    u = "ms"
                    dur = x
                if k == "dram__bytes_read.sum":
                    dr = x
                if k == "dram__bytes_write.sum":
                    dw = x
                f.write("| `%s` | %.6g %s |\n" % (k, x, u))
            if dr is not None and dw is not None and dur:
                f.write("| DRAM read + write / duration | %.0f GB/s |\n" % ((dr + dw) / 1e3 / (dur * 1e-3)))
            f.write("\n")
    print("wrote", out)


main()
