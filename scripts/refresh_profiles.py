"""Rebuild the tracked summaries under profiles/ from one evidence run of scripts/gpu_profile.sh
(gpurun_out/<tag>_*): ncu raw pages -> markdown tables, launch list -> shares, DRAM traffic of the
k-NN pass -> JSON (read by bench.py), the bench line, the SASS opcode histogram.
    python scripts/refresh_profiles.py <tag>"""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
G = os.path.join(ROOT, "gpurun_out", tag)
P = os.path.join(ROOT, "profiles")
run = "run %s" % tag

SUMMARIES = [
    ("knn_pass", "r2: ncu --set full of the k-NN tensor pass (k_knn_2sm<EPI_LISTS, 128>: cta_group::2, folded-norm FP16, pipelined early hand-back drain)",
     "`ncu --set full --clock-control none --import-source on -k regex:k_knn_2sm -s 1 -c 1` over `scripts/prof_knn.py 10000 1000000 128 100 1` (BASELINE configs[1]); final code of round 2 (%s)." % run),
    ("knn_rest", "r2: ncu --set full of the kernels around the k-NN tensor pass",
     "`-k regex:'k_rerank|k_merge_lists|k_row_kth|k_center_rows' -c 5` over the same command (%s)." % run),
    ("kmeans_update", "r2: ncu --set full of the k-means centroid update (BASELINE configs[3]: n = 10M, d = 128, k = 65536)",
     "`-k regex:'k_segsum|k_scatter_ids|k_hist|k_scan_u32|k_seg_counts|k_sum_dis' -c 8` over `scripts/prof_kmeans_update.py 1` (%s)." % run),
    ("kmeans_assign", "r2: ncu --set full of the k-means assignment (k_knn_2sm<EPI_NEAREST, 128>) and its exact re-rank (k_rerank_k1_lanes: lane per (point, candidate) pair)",
     "`-k regex:'k_knn_2sm|k_rerank_k1' -c 2` over `scripts/prof_kmeans.py 1250000 128 65536 1` (one eighth of BASELINE configs[3]: the shard of an 8-GPU run; %s)." % run),
    ("hamming", "r2: ncu --set full of the Hamming tensor engine (BASELINE configs[2]: 10M x 64 bit, 10k queries, k = 100)",
     "`ONLY_TC=1 ... -k regex:'k_knn_tf32|k_knn_2sm|k_ham_tc_finish|k_ham_expand' -c 5` over `scripts/prof_hamming.py`: expansion of the database and the queries, sampling pass (k_knn_2sm<3, 128, 5>), the E4M3 pass (k_knn_2sm<0, 128, 5>: one row per accumulator, constant norm), order + certify (%s)." % run),
    ("hamming_scan", "r2: ncu --set full of the Hamming popcount scan (engine 0, BASELINE configs[2])",
     "`-k regex:k_nn_hamming_scan -c 1` over `scripts/prof_hamming.py` (%s)." % run),
    ("cross", "r2: ncu --set full of compute_cross_distances on the tensor cores (10k x 100k x 128)",
     "`-k regex:'k_knn_2sm|k_split_rows_h' -c 3` over `scripts/prof_cross.py`: the two operand conversions and the K = 3 d contraction (k_knn_2sm<EPI_CROSS, 16, F16N>: wide resident query tile, staged TMA stores; %s)." % run),
]
for name, title, note in SUMMARIES:
    raw = "%s_%s.raw.csv" % (G, name)
    if os.path.exists(raw) and os.path.getsize(raw) > 0:
        subprocess.check_call([sys.executable, os.path.join(ROOT, "scripts", "summarize_raw.py"), raw,
                               os.path.join(P, "r2_%s_ncu.md" % name), title, note])

# launch list
rows = list(csv.reader(open(G + "_launches.csv")))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr, data = rows[hi], rows[hi + 1:]
ci = {n: i for i, n in enumerate(hdr)}
agg = collections.OrderedDict()
for r in data:
    if len(r) < len(hdr):
        continue
    try:
        v = float(r[ci["Metric Value"]].replace(",", ""))
    except ValueError:
        continue
    v *= {"us": 1e-3, "ns": 1e-6, "s": 1e3, "ms": 1.0}.get(r[ci["Metric Unit"]], 1.0)
    a = agg.setdefault(r[ci["Kernel Name"]], [0, 0.0])
    a[0] += 1
    a[1] += v
ours = {k: v for k, v in agg.items() if "yb::" in k}
tot = sum(v[1] for v in ours.values())
with open(os.path.join(P, "r2_launches.md"), "w") as f:
    f.write("# r2: ncu launch list of `python bench.py --steps 2 --warmup 1 --kmeans-iters 2` (%s, final code)\n\n" % run)
    f.write("`ncu --metrics gpu__time_duration.sum --clock-control none -c 6000` (cold-cache, serialised:\n"
            "compare SHARES, not absolutes).  The command runs the k-NN headline, the Hamming, k-means, cross-distance and\n"
            "consumer blocks.  Template arguments of `k_knn_2sm`: <epilogue mode, columns per TMEM load, operand kind, streamed>;\n"
            "modes 0 top-k', 1 k = 1 margin, 3 sampling, 6 cross distances; kinds 3 folded-norm FP16, 5 constant-norm E4M3.\n\n")
    f.write("| kernel | launches | total ms | share of our kernels |\n|---|---:|---:|---:|\n")
    for k, (n, ms) in sorted(ours.items(), key=lambda kv: -kv[1][1]):
        f.write("| `%s` | %d | %.3f | %.1f %% |\n" % (k.split("(")[0], n, ms, 100 * ms / tot))
    f.write("\nOther kernels in the process:\n\n")
    for k, (n, ms) in agg.items():
        if k not in ours:
            f.write("* `%s` x%d, %.3f ms\n" % (k[:100], n, ms))
open(os.path.join(P, "r2_launches.csv"), "w").write(open(G + "_launches.csv").read())

# DRAM traffic of the k-NN pass
rr = list(csv.reader(open(G + "_knn_pass.raw.csv")))
m = {n: (u, v) for n, u, v in zip(rr[0], rr[1], rr[2])}
tb = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def g(n):
    u, v = m[n]
    return float(v.replace(",", "")) * tb.get(u, 1)


commit = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True, cwd=ROOT).stdout.strip()
json.dump({"kernel": "k_knn_2sm<EPI_LISTS, 128> (cta_group::2 folded-norm FP16 pass)",
           "dram_bytes_per_launch": g("dram__bytes_read.sum") + g("dram__bytes_write.sum"), "commit": commit,
           "source": "ncu --set full --clock-control none, gpurun_out/%s_knn_pass.ncu-rep (%s)" % (tag, run)},
          open(os.path.join(P, "r2_knn_tf32_traffic.json"), "w"))
open(os.path.join(P, "r2_bench.json"), "w").write(open(G + "_bench.json").read())
with open(os.path.join(P, "r2_sass_opcodes.txt"), "w") as f:
    subprocess.check_call([sys.executable, os.path.join(ROOT, "scripts", "sass_histogram.py")], stdout=f)
print("pass under ncu:", m["gpu__time_duration.sum"], "tensor pipe",
      m["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"])
