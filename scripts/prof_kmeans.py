"""Time k-means iterations on device-resident points.  Usage:
    python scripts/prof_kmeans.py [n] [d] [k] [niter]"""
import ctypes as C
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import yael_b200
from yael_b200.ynumpy import KMEANS_INIT_USER, KMEANS_QUIET

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
d = int(sys.argv[2]) if len(sys.argv) > 2 else 128
k = int(sys.argv[3]) if len(sys.argv) > 3 else 4096
niter = int(sys.argv[4]) if len(sys.argv) > 4 else 3
L = yael_b200.lib()
torch.manual_seed(1237)
v = torch.rand((n, d), device="cuda", dtype=torch.float32)
cent = v[torch.randperm(n, device="cuda")[:k]].cpu().numpy().copy()
nassign = np.empty(k, np.int32)
f, i = C.POINTER(C.c_float), C.POINTER(C.c_int)
L.yb_prof_enable(1)
for rep in range(2):
    c = cent.copy()
    torch.cuda.synchronize()
    t = time.perf_counter()
    q = L.yb_kmeans_dev(d, n, k, niter, v.data_ptr(), KMEANS_INIT_USER | KMEANS_QUIET, 1, 1,
                        c.ctypes.data_as(f), None, None, nassign.ctypes.data_as(i), None, None)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t
    print("rep %d: %d iterations in %.3f s -> %.3f iter/s, qerr %.5f, engine %d uncert %d" %
          (rep, niter, dt, niter / dt, q, L.yb_last_knn_engine(), L.yb_last_knn_uncertified()))
print("cluster sizes after the run: min %d  median %d  p99 %d  p99.9 %d  max %d  (n/k = %.0f)" % (
    nassign.min(), np.median(nassign), np.percentile(nassign, 99), np.percentile(nassign, 99.9), nassign.max(), n / k))
cnt = C.c_long(0)
for ph in range(20):
    ms = L.yb_prof_ms(ph, C.byref(cnt), 0)
    if cnt.value:
        print("phase %d: %.3f ms avg over %d" % (ph, ms / cnt.value, cnt.value))
