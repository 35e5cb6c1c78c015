"""Turn the ncu artefacts brought back in gpurun_out/ into the small tracked summaries under
profiles/ (the .ncu-rep files themselves are scratch).  Usage:
    python scripts/summarize_profiles.py <round-tag> <launches.csv> <full.ncu-rep> [bench.json]"""
import collections
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

KEYS = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.avg.per_cycle_elapsed",
        "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio"]


def one_kernel(rep, out_md, title, note):
    """--kernel mode: the raw page of ONE captured launch -> a small markdown table."""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(io.StringIO(raw)))
    m = {n: (u, v) for n, u, v in zip(rr[0], rr[1], rr[2])}
    with open(out_md, "w") as f:
        f.write("# %s\n\n%s\n\n| metric | value | unit |\n|---|---:|---|\n" % (title, note))
        f.write("| kernel | `%s` | |\n" % m.get("Kernel Name", ("", "?"))[1][:90])
        for k in KEYS:
            u, v = m.get(k, ("", "nan"))
            f.write("| `%s` | %s | %s |\n" % (k, v, u))
    print("wrote", out_md)


if sys.argv[1] == "--kernel":  # summarize_profiles.py --kernel <rep> <out.md> <title> <note>
    one_kernel(sys.argv[2], sys.argv[3], sys.argv[4], sys.argv[5] if len(sys.argv) > 5 else "")
    sys.exit(0)

tag, launches, rep = sys.argv[1], sys.argv[2], sys.argv[3]
bench = sys.argv[4] if len(sys.argv) > 4 else None
out = os.path.join(ROOT, "profiles")
os.makedirs(out, exist_ok=True)

# ---- launch list: per-kernel totals and shares
rows = list(csv.reader(open(launches)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr, data = rows[hi], rows[hi + 1:]
ci = {n: i for i, n in enumerate(hdr)}
agg = collections.OrderedDict()
for r in data:
    if len(r) < len(hdr):
        continue
    try:
        v = float(r[ci["Metric Value"]])
    except ValueError:
        continue
    v *= {"us": 1e-3, "ns": 1e-6, "s": 1e3, "ms": 1.0}.get(r[ci["Metric Unit"]], 1.0)
    a = agg.setdefault(r[ci["Kernel Name"]], [0, 0.0])
    a[0] += 1
    a[1] += v
ours = {k: v for k, v in agg.items() if k.startswith("yb::") or "yb::" in k}
tot_ours = sum(v[1] for v in ours.values())
with open(os.path.join(out, "%s_launches.md" % tag), "w") as f:
    f.write("# %s: ncu launch list of `python bench.py --steps 2 --warmup 1`\n\n" % tag)
    f.write("`ncu --metrics gpu__time_duration.sum --clock-control none -c 400` (cold-cache, serialised:\n"
            "compare SHARES, not absolutes).  Kernels of this library only; the cuBLAS TF32 GEMM of the\n"
            "in-run peak measurement and torch's RNG kernels are listed at the end.\n\n")
    f.write("| kernel | launches | total ms | share of our kernels |\n|---|---:|---:|---:|\n")
    for k, (n, ms) in sorted(ours.items(), key=lambda kv: -kv[1][1]):
        f.write("| `%s` | %d | %.3f | %.1f %% |\n" % (k.split("(")[0], n, ms, 100 * ms / tot_ours))
    f.write("\nOther kernels in the process:\n\n")
    for k, (n, ms) in agg.items():
        if k not in ours:
            f.write("* `%s` x%d, %.3f ms\n" % (k[:100], n, ms))
open(os.path.join(out, "%s_launches.csv" % tag), "w").write(open(launches).read())

# ---- full capture of the top kernel
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(io.StringIO(raw)))
names, units, vals = rr[0], rr[1], rr[2]
m = {n: (u, v) for n, u, v in zip(names, units, vals)}


def g(name):
    u, v = m.get(name, ("", "nan"))
    try:
        return float(v.replace(",", "")), u
    except ValueError:
        return float("nan"), u


keys = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.avg.per_cycle_elapsed",
        "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum",
        "lts__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"]
to_bytes = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
dr, dru = g("dram__bytes_read.sum")
dw, dwu = g("dram__bytes_write.sum")
traffic = dr * to_bytes.get(dru, 1) + dw * to_bytes.get(dwu, 1)
json.dump({"kernel": "k_knn_tf32", "dram_bytes_per_launch": traffic,
           "source": "ncu --set full --clock-control none, %s" % os.path.basename(rep)},
          open(os.path.join(out, "%s_knn_tf32_traffic.json" % tag), "w"))
with open(os.path.join(out, "%s_knn_tf32_ncu.md" % tag), "w") as f:
    f.write("# %s: `ncu --set full --clock-control none --import-source on -k regex:k_knn_tf32`\n\n" % tag)
    f.write("Full pass of the bench workload (10 000 queries x 1 000 000 rows x 128, k'=200).\n"
            "Durations under the profiler are not bench values.\n\n| metric | value | unit |\n|---|---:|---|\n")
    for k in keys:
        v, u = g(k)
        f.write("| `%s` | %.6g | %s |\n" % (k, v, u))
    f.write("\nDRAM traffic per launch: %.3f GB (algorithmic minimum: database 0.512 GB + queries + lists)\n" % (traffic / 1e9))
if bench:
    open(os.path.join(out, "%s_bench.json" % tag), "w").write(open(bench).read())
print("wrote", sorted(os.listdir(out)))
