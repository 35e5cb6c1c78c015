"""compute_cross_distances at 10 000 x 100 000 x 128 on device-resident matrices, both engines."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import yael_b200

L = yael_b200.lib()
na, nb, d = 10000, 100000, 128
torch.manual_seed(4242)
a = torch.rand((na, d), device="cuda")
b = torch.rand((nb, d), device="cuda")
out = torch.empty((nb, na), device="cuda")
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
