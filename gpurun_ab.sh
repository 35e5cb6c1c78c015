python -m pytest tests/test_gpu_tf32_path.py tests/test_gpu_golden_and_merge.py tests/test_gpu_exact_path.py -x -q -m gpu 2>&1 | tail -3
YAEL_B200_TWO_LEVEL=1 python scripts/prof_knn.py 10000 1000000 128 100 6 2>&1 | tail -7
python scripts/prof_knn.py 10000 1000000 128 100 6 2>&1 | tail -7
python scripts/prof_knn.py 10000 1000000 128 10 4 2>&1 | tail -7
python scripts/prof_knn.py 10000 1000000 128 1000 4 2>&1 | tail -7
python scripts/prof_knn.py 2000 4000000 96 100 4 2>&1 | tail -7
