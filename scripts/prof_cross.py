"""compute_cross_distances at 10 000 x 100 000 x 128 on device-resident matrices, both engines."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import yael_b200

L = yael_b200.lib()
na = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
nb = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
d = int(sys.argv[3]) if len(sys.argv) > 3 else 128
torch.manual_seed(4242)
a = torch.rand((na, d), device="cuda")
b = torch.rand((nb, d), device="cuda")
out = torch.empty((nb, na), device="cuda")
L.yb_prof_enable(1)
for engine in (1, 0):
    L.yb_set_cross_engine(engine)
    for rep in range(3):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = L.yb_cross_distances_l2(d, na, nb, a.data_ptr(), d, b.data_ptr(), d, out.data_ptr(), na, C.c_void_p(1))
        e1.record()
        torch.cuda.synchronize()
        assert rc == 0, L.yb_last_error()
        print("engine %d (used %d) rep %d: %.3f ms" % (engine, L.yb_last_cross_engine(), rep, e0.elapsed_time(e1)))
L.yb_set_cross_engine(-1)
cnt = C.c_long(0)
for ph in range(12):
    ms = L.yb_prof_ms(ph, C.byref(cnt), 0)
    if cnt.value:
        print("phase %d: %.3f ms avg over %d" % (ph, ms / cnt.value, cnt.value))
if int(os.environ.get("YAEL_B200_TF32_DEBUG", "0")) & 512:
    import numpy as np
    ck = np.zeros((148, 16), np.int64)
    L.yb_debug_tf32_clocks(ck.ctypes.data_as(C.c_void_p), 148)
    print("clock attribution (per CTA totals, mean over CTAs with data):")
    names = ["issuer: wait accumulator", "issuer: wait operands", "issuer: wait extras", "issuer: total",
             "epilogue: wait accumulator", "epilogue: drain", "epilogue: hand back", "epilogue: total", "tiles"]
    for i, nme in enumerate(names):
        sel = ck[:, i] > 0
        print("  %-28s %12.0f  (%d CTAs)" % (nme, ck[sel, i].mean() if sel.any() else 0.0, sel.sum()))
