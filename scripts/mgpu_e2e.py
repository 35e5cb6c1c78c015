"""The drop-in calls of ONE process on host buffers, sharded over the box's GPUs by the library
itself (yb_mgpu.cu) against the same calls pinned to one GPU: knn_full_thread at BASELINE configs[1],
nn_hamming at configs[2], kmeans at a reduced configs[3].  Prints one JSON line.
Usage: python scripts/mgpu_e2e.py [reps]"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # pinned host buffers only

import yael_b200
from yael_b200 import ynumpy

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
L = yael_b200.lib()
ndev = L.yb_mgpu_device_count()
out = {"gpus": ndev}


def pinned(a):
    t = torch.from_numpy(a).pin_memory()
    return t.numpy(), t


def timed(fn):
    fn()
    best = 1e9
    for _ in range(reps):
        t = time.perf_counter()
        r = fn()
        best = min(best, time.perf_counter() - t)
    return best, r


one = (C.c_int * 1)(0)
r = np.random.RandomState(1234)
base, _b = pinned(r.random_sample((1_000_000, 128)).astype(np.float32))
query, _q = pinned(np.random.RandomState(1235).random_sample((10_000, 128)).astype(np.float32))
L.yb_mgpu_set_devices(0, None)
t_all, (ia, da) = timed(lambda: ynumpy.knn(query, base, 100))
used = L.yb_mgpu_last_used()
L.yb_mgpu_set_devices(1, one)
t_one, (i1, d1) = timed(lambda: ynumpy.knn(query, base, 100))
out["knn_full_1Mx128_10kq_k100"] = {"ms_all_gpus": t_all * 1e3, "gpus_used": used, "ms_one_gpu": t_one * 1e3,
                                     "identical": bool(np.array_equal(ia, i1) and np.array_equal(da, d1))}

codes, _c = pinned(r.randint(0, 256, (10_000_000, 8)).astype(np.uint8))
qc, _qc = pinned(r.randint(0, 256, (10_000, 8)).astype(np.uint8))
L.yb_mgpu_set_devices(0, None)
t_all, (ia, da) = timed(lambda: ynumpy.knn_hamming(qc, codes, 100))
used = L.yb_mgpu_last_used()
L.yb_mgpu_set_devices(1, one)
t_one, (i1, d1) = timed(lambda: ynumpy.knn_hamming(qc, codes, 100))
out["nn_hamming_10Mx64bit_10kq_k100"] = {"ms_all_gpus": t_all * 1e3, "gpus_used": used, "ms_one_gpu": t_one * 1e3,
                                          "identical": bool(np.array_equal(ia, i1) and np.array_equal(da, d1))}

n, d, k, niter = 2_000_000, 128, 16384, 3
v, _v = pinned(r.random_sample((n, d)).astype(np.float32))
L.yb_mgpu_set_devices(0, None)
t_all, ra = timed(lambda: ynumpy.kmeans(v, k, niter=niter, verbose=False, seed=11, output="all"))
used = L.yb_mgpu_last_used()
L.yb_mgpu_set_devices(1, one)
t_one, r1 = timed(lambda: ynumpy.kmeans(v, k, niter=niter, verbose=False, seed=11, output="all"))
out["kmeans_2Mx128_k16384_3it"] = {"ms_all_gpus": t_all * 1e3, "gpus_used": used, "ms_one_gpu": t_one * 1e3,
                                    "qerr": [float(ra[1]), float(r1[1])],
                                    "assign_differing": int((ra[3] != r1[3]).sum()),
                                    "max_abs_centroid_diff": float(np.abs(ra[0] - r1[0]).max())}
L.yb_mgpu_set_devices(0, None)
print(json.dumps(out))
