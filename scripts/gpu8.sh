#!/bin/bash
# 8-GPU evidence run (under gpurun --gpus 8): 2-GPU parity tests, the bench line at N = 8 (with the
# sharded k-means / Hamming blocks and the round-1 exchange as A/B), BASELINE configs[4] (100M x 96
# sharded), and the drop-in calls of ONE process sharded by the library itself.
tag=${1:-rX}
N=${2:-8}
out=gpurun_out
mkdir -p $out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 python -m pytest tests/test_gpu_multi.py -x -q > $out/${tag}_multi.log 2>&1; tail -3 $out/${tag}_multi.log
timeout 600 $TR --master-port 29751 bench.py --gpus $N --steps 10 --warmup 3 > $out/${tag}_bench$N.json 2> $out/${tag}_bench$N.err
tail -c 400 $out/${tag}_bench$N.json; tail -2 $out/${tag}_bench$N.err
timeout 600 $TR --master-port 29752 bench.py --gpus $N --workload c5 --steps 3 --warmup 2 > $out/${tag}_c5_$N.json 2> $out/${tag}_c5_$N.err
tail -c 600 $out/${tag}_c5_$N.json; tail -2 $out/${tag}_c5_$N.err
timeout 400 python scripts/mgpu_e2e.py 3 > $out/${tag}_mgpu$N.json 2> $out/${tag}_mgpu$N.err
tail -c 1200 $out/${tag}_mgpu$N.json; tail -2 $out/${tag}_mgpu$N.err
