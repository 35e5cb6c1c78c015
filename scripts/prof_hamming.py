"""Time nn_hamming on device-resident codes.  Usage:
    python scripts/prof_hamming.py [nq] [nb] [ncodes] [k]"""
import ctypes as C
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import yael_b200

nq = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
nb = int(sys.argv[2]) if len(sys.argv) > 2 else 10000000
nc = int(sys.argv[3]) if len(sys.argv) > 3 else 8
k = int(sys.argv[4]) if len(sys.argv) > 4 else 100
L = yael_b200.lib()
torch.manual_seed(1236)
base = torch.randint(0, 256, (nb, nc), device="cuda", dtype=torch.uint8)
query = torch.randint(0, 256, (nq, nc), device="cuda", dtype=torch.uint8)
idx = torch.empty((nq, k), device="cuda", dtype=torch.int32)
dis = torch.empty((nq, k), device="cuda", dtype=torch.int16)
engines = (1,) if os.environ.get("ONLY_TC") else (0, 1)
for engine in engines:
    L.yb_set_hamming_engine(engine)
    L.yb_prof_enable(1)
    L.yb_prof_ms(0, None, 1)
    for rep in range(3):
        torch.cuda.synchronize()
        t = time.perf_counter()
        rc = L.yb_nn_hamming(nq, nb, nc, k, base.data_ptr(), query.data_ptr(), idx.data_ptr(), dis.data_ptr(), 0, None)
        assert rc == 0, L.yb_last_error()
        L.yb_sync(None)
        dt = time.perf_counter() - t
        print("engine %d (used %d, %d scan fallbacks) rep %d: %.3f ms -> %.0f q/s, %.3e pairs/s" %
              (engine, L.yb_last_hamming_engine(), L.yb_last_hamming_fallbacks(), rep, dt * 1e3, nq / dt,
               nq * nb / dt))
    cnt = C.c_long(0)
    for ph, name in ((7, "popcount scan"), (12, "expand"), (13, "sample"), (14, "e4m3 pass"), (15, "order+certify")):
        ms = L.yb_prof_ms(ph, C.byref(cnt), 0)
        if cnt.value:
            print("  phase %-14s %.3f ms avg over %d" % (name, ms / cnt.value, cnt.value))
    L.yb_prof_enable(0)
    print(idx[0, :5].tolist(), dis[0, :5].tolist())
    if int(os.environ.get("YAEL_B200_TF32_DEBUG", "0")) & 512 and engine == 1:
        import numpy as np
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
    if engine == 0:
        ref = (idx.clone(), dis.clone())
    elif len(engines) > 1:
        print("engines agree:", bool(torch.equal(ref[0], idx) and torch.equal(ref[1], dis)))
L.yb_set_hamming_engine(-1)
