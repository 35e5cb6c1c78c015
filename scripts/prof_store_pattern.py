"""Store-pattern micro-benchmark behind the compute_cross_distances epilogue (yb_debug_store_pattern_gbs)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import yael_b200

L = yael_b200.lib()
nq, nb = 10000, 100000
for ld in (10000, 10016, 10240):
    out = torch.empty((nb, ld), device="cuda")
    for mode in (0, 1, 2):
        gbs = L.yb_debug_store_pattern_gbs(out.data_ptr(), ld, nq, nb, mode, None)
        print("ld %5d mode %d: %.0f GB/s" % (ld, mode, gbs))
    del out
