"""hkm_quantize (1M x 128, bf 10, 3 levels) and gmm_compute_p (200k x 64 x 256) on device-resident
points, for ncu captures of k_gmm_logp / k_gmm_softmax / k_hkm_* / k_rerank_k1_lanes."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import yael_b200
from yael_b200 import _lib as yl

L = yael_b200.lib()
f = C.POINTER(C.c_float)
r = np.random.RandomState(77)
dev = torch.device("cuda", 0)
n, d, bf, nl = 1_000_000, 128, 10, 3
levels = [r.rand(bf ** (l + 1), d).astype(np.float32) for l in range(nl)]
ptrs = (f * nl)(*[x.ctypes.data_as(f) for x in levels])
h = yl.HkmT(nl, bf, bf ** nl, d, C.cast(ptrs, C.POINTER(f)))
v = torch.rand((n, d), device=dev)
idx = torch.empty(n, dtype=torch.int32, device=dev)
for _ in range(2):
    L.hkm_quantize(C.byref(h), n, C.cast(v.data_ptr(), f), C.cast(idx.data_ptr(), C.POINTER(C.c_int)))
torch.cuda.synchronize()
n, d, k = 200_000, 64, 256
mu = r.rand(k, d).astype(np.float32)
sg = (0.05 + 0.2 * r.rand(k, d)).astype(np.float32)
w = np.full(k, 1.0 / k, np.float32)
g = yl.GmmT(d, k, w.ctypes.data_as(f), mu.ctypes.data_as(f), sg.ctypes.data_as(f))
v2 = torch.rand((n, d), device=dev)
p = torch.empty((n, k), device=dev)
for _ in range(2):
    L.gmm_compute_p(n, C.cast(v2.data_ptr(), f), C.byref(g), C.cast(p.data_ptr(), f), 1)
torch.cuda.synchronize()
print("done")
