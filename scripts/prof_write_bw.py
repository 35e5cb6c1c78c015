"""Pure-write and copy bandwidth of the box (torch fill_ / copy_), the yardstick for output-bound kernels."""
import torch

def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

for gb in (0.1, 1, 4):
    n = int(gb * 1e9 / 4)
    a = torch.empty(n, device="cuda")
    b = torch.empty(n, device="cuda")
    t = timed(lambda: a.fill_(1.0))
    print("fill  %4.1f GB: %.3f ms  %.0f GB/s written" % (gb, t, 4 * n / t / 1e6))
    t = timed(lambda: a.zero_())
    print("zero  %4.1f GB: %.3f ms  %.0f GB/s written" % (gb, t, 4 * n / t / 1e6))
    t = timed(lambda: b.copy_(a))
    print("copy  %4.1f GB: %.3f ms  %.0f GB/s read + written" % (gb, t, 8 * n / t / 1e6))
    del a, b
