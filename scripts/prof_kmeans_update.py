"""Times yb_kmeans_accumulate at the BASELINE configs[3] shape (n = 10M, d = 128, k = 65536) with a
synthetic uniform assignment: both update paths, CUDA events on the launching stream.
Run under ncu for the per-kernel list."""
import ctypes as C
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import yael_b200  # noqa: E402

L = yael_b200.lib()
dev = torch.device("cuda", 0)
L.yb_set_device(0)
n, d, k = 10_000_000, 128, 65536
torch.manual_seed(0)
v = torch.rand((n, d), device=dev)
assign = torch.randint(0, k, (n,), device=dev, dtype=torch.int32)
dis = torch.rand(n, device=dev)
sums = torch.empty((k, d), device=dev)
cnt = torch.empty(k, device=dev, dtype=torch.int32)
q = torch.empty(1, device=dev, dtype=torch.float64)
st = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(st)
sp = C.c_void_p(st.cuda_stream)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
for mode in ("fast", "general"):
    if mode == "general":
        os.environ["YAEL_B200_KMEANS_GENERAL_UPDATE"] = "1"
    for _ in range(2):
        L.yb_kmeans_accumulate(d, n, k, v.data_ptr(), assign.data_ptr(), dis.data_ptr(), sums.data_ptr(),
                               cnt.data_ptr(), q.data_ptr(), 0, sp)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(reps):
        L.yb_kmeans_accumulate(d, n, k, v.data_ptr(), assign.data_ptr(), dis.data_ptr(), sums.data_ptr(),
                               cnt.data_ptr(), q.data_ptr(), 0, sp)
    e1.record(st)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    gb = (4.0 * n * d + 4.0 * n + 4.0 * k * d + 4.0 * k) / 1e9
    print("%s update: %.3f ms  %.0f GB/s algorithmic (%.2f GB)" % (mode, ms, gb / (ms * 1e-3), gb), flush=True)
    ref = sums.clone() if mode == "fast" else ref
    if mode == "general":
        print("paths agree bit for bit:", bool(torch.equal(ref, sums)))
