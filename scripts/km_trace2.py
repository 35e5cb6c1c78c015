"""Sharded k-means (BASELINE configs[3] split over the ranks) with the host-side timeline switched on."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import yael_b200
from yael_b200 import dist as ydist
world = int(os.environ["WORLD_SIZE"]); rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
L = yael_b200.lib(); L.yb_set_device(local)
dist.init_process_group("nccl", device_id=dev)
n, d, k = 10_000_000, 128, 65536
lo, hi = ydist.shard_bounds(n, world)[rank]
g = torch.Generator(device=dev); g.manual_seed(1237 + rank)
v = torch.rand((hi - lo, d), device=dev, generator=g)
cent = v[:k].contiguous() if rank == 0 else torch.empty((k, d), device=dev)
dist.broadcast(cent, 0)
c0 = cent.cpu().numpy()
ydist.sharded_kmeans(v, k, 2, c0, n)
os.environ["YAEL_B200_KM_TRACE"] = "1"
dist.barrier(); torch.cuda.synchronize()
t = time.perf_counter()
ydist.sharded_kmeans(v, k, 4, c0, n)
torch.cuda.synchronize()
if rank == 0:
    print("s per iteration: %.4f" % ((time.perf_counter() - t) / 4))
dist.destroy_process_group()
