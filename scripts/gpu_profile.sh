#!/bin/bash
# One-GPU evidence run (under gpurun): GPU tests, the bench line, the ncu launch list of the bench
# command, and `ncu --set full` captures of the dominant kernels.  Everything lands in gpurun_out/
# with the tag given as $1; the large .ncu-rep files are reduced to raw-page CSVs on the box
# (only the k-NN tensor pass keeps its report, for the source page).
tag=${1:-rX}
out=gpurun_out
mkdir -p $out
NCU="ncu --clock-control none"
if [ -z "$SKIP_TESTS" ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1
  tail -3 $out/${tag}_pytest.log
fi
timeout 900 python bench.py --steps 20 --warmup 3 > $out/${tag}_bench.json 2> $out/${tag}_bench.err
tail -c 600 $out/${tag}_bench.json
timeout 900 $NCU --metrics gpu__time_duration.sum -c 6000 --csv --log-file $out/${tag}_launches.csv \
  python bench.py --steps 2 --warmup 1 --kmeans-iters 2 > $out/${tag}_ncu_bench.log 2>&1
# the k-NN tensor pass (with source) and the kernels around it
timeout 600 $NCU --set full --import-source on -k regex:k_knn_2sm -s 1 -c 1 -f -o $out/${tag}_knn_pass \
  python scripts/prof_knn.py 10000 1000000 128 100 1 > $out/${tag}_ncu_knn.log 2>&1
timeout 600 $NCU --set full -k regex:'k_rerank|k_merge_lists|k_row_kth|k_center_rows' -c 5 -f -o $out/${tag}_knn_rest \
  python scripts/prof_knn.py 10000 1000000 128 100 1 >> $out/${tag}_ncu_knn.log 2>&1
# the k-means centroid update at the BASELINE configs[3] shape
timeout 600 $NCU --set full -k regex:'k_segsum|k_scatter_ids|k_hist|k_scan_u32|k_seg_counts|k_sum_dis' -c 8 -f \
  -o $out/${tag}_kmeans_update python scripts/prof_kmeans_update.py 1 > $out/${tag}_ncu_kmeans.log 2>&1
# the k-means assignment (k = 1 margin mode) and its exact re-rank at one eighth of BASELINE configs[3]
timeout 600 $NCU --set full -k regex:'k_knn_2sm|k_rerank_k1' -c 2 -f \
  -o $out/${tag}_kmeans_assign python scripts/prof_kmeans.py 1250000 128 65536 1 > $out/${tag}_ncu_kmeans_assign.log 2>&1
# Hamming: the tensor engine (expansion, sampling pass, the E4M3 pass, order + certify), then the popcount scan
ONLY_TC=1 timeout 900 $NCU --set full -k regex:'k_knn_tf32|k_knn_2sm|k_ham_tc_finish|k_ham_expand' -c 5 -f \
  -o $out/${tag}_hamming python scripts/prof_hamming.py > $out/${tag}_ncu_hamming.log 2>&1
YAEL_B200_HAMMING_ENGINE=0 timeout 900 $NCU --set full -k regex:'k_nn_hamming_scan' -c 1 -f \
  -o $out/${tag}_hamming_scan python scripts/prof_hamming.py >> $out/${tag}_ncu_hamming.log 2>&1
# compute_cross_distances on the tensor cores
timeout 600 $NCU --set full -k regex:'k_knn_2sm|k_split_rows_h' -c 3 -f \
  -o $out/${tag}_cross python scripts/prof_cross.py > $out/${tag}_ncu_cross.log 2>&1
for r in knn_pass knn_rest kmeans_update kmeans_assign hamming hamming_scan cross; do
  f=$out/${tag}_$r.ncu-rep
  [ -f $f ] && ncu -i $f --page raw --csv > $out/${tag}_$r.raw.csv 2>/dev/null
done
rm -f $out/${tag}_knn_rest.ncu-rep $out/${tag}_kmeans_update.ncu-rep $out/${tag}_kmeans_assign.ncu-rep \
  $out/${tag}_hamming.ncu-rep $out/${tag}_hamming_scan.ncu-rep $out/${tag}_cross.ncu-rep
ls -la $out | tail -20
