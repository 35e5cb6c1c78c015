"""Run the C2-shaped kNN a few times (for ncu / launch lists).  Usage:
    python scripts/prof_knn.py [nq] [nb] [d] [k] [reps]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np

import yael_b200
from devmem import DevArray

nq = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
nb = int(sys.argv[2]) if len(sys.argv) > 2 else 1000000
d = int(sys.argv[3]) if len(sys.argv) > 3 else 128
k = int(sys.argv[4]) if len(sys.argv) > 4 else 100
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 3
L = yael_b200.lib()
r = np.random.RandomState(1234)
base = DevArray(r.random_sample((nb, d)).astype(np.float32))
query = DevArray(r.random_sample((nq, d)).astype(np.float32))
idx = DevArray(shape=(nq, k), dtype=np.int32)
dis = DevArray(shape=(nq, k), dtype=np.float32)
L.yb_prof_enable(1)
for i in range(reps):
    t = time.perf_counter()
    rc = L.yb_knn_l2(nq, nb, d, k, base.ptr, query.ptr, None, idx.ptr, dis.ptr, 0, None)
    assert rc == 0, L.yb_last_error()
    L.yb_sync(None)
    print("rep %d: %.3f ms engine %d uncert %d" % (i, (time.perf_counter() - t) * 1e3,
                                                   L.yb_last_knn_engine(), L.yb_last_knn_uncertified()))
import ctypes as C
cnt = C.c_long(0)
for ph in range(12):
    ms = L.yb_prof_ms(ph, C.byref(cnt), 0)
    if cnt.value:
        print("phase %d: %.3f ms avg over %d" % (ph, ms / cnt.value, cnt.value))

if int(os.environ.get("YAEL_B200_TF32_DEBUG", "0")) & 512:
    ck = np.zeros((148, 16), np.int64)
    L.yb_debug_tf32_clocks(ck.ctypes.data_as(C.c_void_p), 148)
    m = ck[ck[:, 8] > 0]
    t = m[:, 8].astype(np.float64)
    names = ["issuer: wait accumulator", "issuer: wait operands", "issuer: wait extras", "issuer: total",
             "epilogue: wait accumulator", "epilogue: drain", "epilogue: hand back", "epilogue: total"]
    print("clock attribution (cycles per tile, mean over %d CTAs, %.0f tiles per CTA; LAST instrumented pass):" % (len(m), t.mean()))
    for i, nme in enumerate(names):
        sel = m[:, i] > 0
        print("  %-28s %8.1f" % (nme, (m[sel, i] / t[sel]).mean() if sel.any() else 0.0))
