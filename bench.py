#!/usr/bin/env python
"""bench.py -- headline benchmark of the yael hot path on B200.

    python bench.py --gpus N --steps K --warmup W              # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...    # the reference's CPU path

Workload (BASELINE.json configs[1]): exact k-NN, SIFT1M shape -- 1M x 128 float32 database,
10 000 queries, k = 100, squared L2, through knn_full's semantics.  One "step" = one pass of
the hot path over the batch of 10 000 queries.  Synthetic uniform[0,1) data, fixed seeds.

  value   queries/s with the database and the queries already resident in HBM (device-level
          C ABI yb_knn_l2 on the caller's stream), CUDA events, max over ranks.
  e2e     the same through the drop-in call knn_full_thread() with HOST (pinned) buffers:
          host->device copies of base + queries and device->host copies of ids + distances
          are inside the timed region.
  roofline  the tcgen05 TF32 shortlist kernel: algorithmic 2*nq*nb*d FLOP / its own average
          duration (CUDA events on the launching stream, yb_prof_*).
  cpu_baseline  the unmodified reference (oracle/_ref, OpenBLAS + OpenMP) or the oracle port on
          the host cores, on a bounded sample of the same workload.

N > 1 (torchrun, one rank per GPU): every rank holds its own 1M x 128 shard of an N-million
row database and scans it for all queries; the per-rank top-k lists are all-gathered over NCCL
and merged on every rank.  value = N * nq / time: query x 1M-shard scans per
second ("weak": per-GPU work is fixed).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NB, NQ, D, K = 1_000_000, 10_000, 128, 100
METRIC = "kNN queries/s (1M x 128 db, k=100)"
UNIT = "queries/s"
WORKLOAD = ("exact kNN, SIFT1M shape: 1M x 128 float32 database per GPU, 10000 queries, k=100 "
            "(BASELINE configs[1]); knn_full semantics")


def gen_data(rank):
    r = np.random.RandomState(1234 + 7919 * rank)
    base = r.random_sample((NB, D)).astype(np.float32)
    rq = np.random.RandomState(1235)  # same queries on every rank
    query = rq.random_sample((NQ, D)).astype(np.float32)
    return base, query


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    """SM clocks / throttle reasons DURING the timed region (B200_PROFILING.md).  NVML is polled
    in-process every 2 ms (a timed region of a few tens of milliseconds sees dozens of samples;
    an `nvidia-smi -lms 100` child, the fallback, may see none).  Samples are time-stamped and
    only those taken between mark_begin() and mark_end() are reported."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    REASONS = (("hw_slowdown", 0x8), ("sw_thermal_slowdown", 0x20), ("hw_thermal_slowdown", 0x40),
               ("sw_power_cap", 0x4))

    def __init__(self, gpu_index, uuid=None):
        self.gpu = gpu_index
        self.uuid = uuid
        self.samples = []   # (t, sm_mhz, sm_max_mhz, reason_bits)
        self.lines = []
        self.proc = None
        self.nvml = None
        self.stop_flag = False
        self.t0 = self.t1 = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = None
            if self.uuid:
                try:
                    h = pynvml.nvmlDeviceGetHandleByUUID(self.uuid)
                except Exception:
                    h = None
            if h is None:
                vis = os.environ.get("CUDA_VISIBLE_DEVICES")
                idx = self.gpu
                if vis:
                    try:
                        idx = int(vis.split(",")[self.gpu])
                    except Exception:
                        idx = self.gpu
                h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.nvml, self.h = pynvml, h
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            self.th = threading.Thread(target=self._poll, daemon=True)
            self.th.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _poll(self):
        n = self.nvml
        while not self.stop_flag:
            try:
                sm = float(n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM))
                try:
                    bits = int(n.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    bits = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                self.samples.append((time.perf_counter(), sm, self.mx, bits))
            except Exception:
                pass
            time.sleep(0.002)

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def mark_begin(self):
        self.t0 = time.perf_counter()

    def mark_end(self):
        self.t1 = time.perf_counter()

    def stop(self):
        if self.nvml is not None:
            self.stop_flag = True
            self.th.join(timeout=1)
            sel = [x for x in self.samples
                   if (self.t0 is None or x[0] >= self.t0) and (self.t1 is None or x[0] <= self.t1)]
            if not sel:
                return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
            reasons = sorted({name for name, bit in self.REASONS for x in sel if x[3] & bit})
            return {"sm_mhz": float(np.median([x[1] for x in sel])), "sm_max_mhz": float(sel[0][2]),
                    "reasons": reasons, "samples": len(sel), "source": "nvml, 2 ms polling"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)),
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi -lms 100"}


# ----------------------------------------------------------------------------- CPU baseline
def cpu_reference_rate(base, query, target_s=15.0, max_q=NQ):
    """Time the reference's own CPU implementation (knn_full_thread, yael/nn.c:679-699) on a
    bounded sample: all host threads, OPENBLAS_NUM_THREADS=1 (yael threads itself)."""
    from oracle import bindings as ob
    cores = len(os.sched_getaffinity(0))
    if ob.have_ref():
        kind = "reference"
        fn = lambda b, q: ob.ref_knn(b, q, K, nt=cores)
    else:
        kind = "port"
        fn = lambda b, q: ob.orc_knn(b, q, K, dot_mode=ob.DOT_F32_SEQ, nt=cores)
    probe = min(4 * cores if 4 * cores >= 64 else 64, max_q)
    t = time.perf_counter()
    fn(base, query[:probe])
    dt = time.perf_counter() - t
    rate = probe / dt
    n = int(min(max_q, max(probe, rate * target_s)))
    n = max(cores, (n // cores) * cores)
    t = time.perf_counter()
    fn(base, query[:n])
    dt = time.perf_counter() - t
    return {"value": n / dt, "unit": UNIT, "cores": cores, "kind": kind,
            "sample": "%d of %d queries against the full 1M x 128 database, k=100, %.1f s" % (n, NQ, dt)}


def run_reference(args):
    """--impl reference: the reference's CPU path on the host cores, same metric / config."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    base, query = gen_data(0)
    from oracle import bindings as ob
    cores = len(os.sched_getaffinity(0))
    kind = "reference" if ob.have_ref() else "port"
    fn = (lambda b, q: ob.ref_knn(b, q, K, nt=cores)) if ob.have_ref() else \
        (lambda b, q: ob.orc_knn(b, q, K, nt=cores))
    # size a step so the whole run (warmup + steps) stays within a few minutes
    probe = min(max(64, 4 * cores), NQ)
    t = time.perf_counter()
    fn(base, query[:probe])
    rate = probe / (time.perf_counter() - t)
    budget = 150.0 / max(1, args.steps + args.warmup)
    n = int(min(NQ, max(probe, rate * budget)))
    n = max(cores, (n // cores) * cores)
    for _ in range(args.warmup):
        fn(base, query[:n])
    t = time.perf_counter()
    for _ in range(args.steps):
        fn(base, query[:n])
    dt = (time.perf_counter() - t) / args.steps
    val = n / dt
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic uniform[0,1), seeds 1234/1235",
        "config": {"workload": WORKLOAD,
                   "step": "%d of the 10000 queries per step (bounded CPU sample), all %d host threads" % (n, cores)},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": "%d queries per step against the full database" % n},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# ----------------------------------------------------------------------------- GPU arm
def measure_tf32_peak(torch, dev):
    """cuBLAS TF32 GEMM throughput measured the way MEASURED_PEAKS.json measures bf16
    (8192^3, best of 10): the denominator for the kind::tf32 kernel."""
    try:
        old = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = True
        n = 8192
        a = torch.randn(n, n, device=dev)
        b = torch.randn(n, n, device=dev)
        for _ in range(3):
            a @ b
        best = 1e9
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            a @ b
            e1.record()
            e1.synchronize()
            best = min(best, e0.elapsed_time(e1))
        torch.backends.cuda.matmul.allow_tf32 = old
        del a, b
        torch.cuda.empty_cache()
        return 2.0 * n ** 3 / (best * 1e-3) / 1e12
    except Exception:
        return None


def measure_extras(torch, L, dev):
    """The other two BASELINE shapes, device-resident, one GPU (reported beside the headline;
    the headline line above stays the kNN workload):
      - k-means, BASELINE configs[3]: n = 10M, d = 128, k = 65536 (iterations/s, 2 iterations)
      - Hamming kNN, BASELINE configs[2]: 10M x 64-bit codes, 10k queries, k = 100"""
    out = {}
    try:
        del_later = []
        torch.manual_seed(1236)
        nbh, nqh = 10_000_000, 10_000
        hb = torch.randint(0, 256, (nbh, 8), device=dev, dtype=torch.uint8)
        hq = torch.randint(0, 256, (nqh, 8), device=dev, dtype=torch.uint8)
        hi = torch.empty((nqh, 100), device=dev, dtype=torch.int32)
        hd = torch.empty((nqh, 100), device=dev, dtype=torch.int16)
        sp = C.c_void_p(torch.cuda.current_stream().cuda_stream)

        def time_engine(engine, reps):
            L.yb_set_hamming_engine(engine)
            best = 1e9
            for _ in range(reps):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                rc = L.yb_nn_hamming(nqh, nbh, 8, 100, hb.data_ptr(), hq.data_ptr(), hi.data_ptr(), hd.data_ptr(), 0, sp)
                e1.record()
                e1.synchronize()
                if rc == 0:
                    best = min(best, e0.elapsed_time(e1))
            used, fb = L.yb_last_hamming_engine(), L.yb_last_hamming_fallbacks()
            L.yb_set_hamming_engine(-1)
            return best, used, fb

        peak_pairs = L.yb_debug_popc_pairs_per_s(sp)
        if not peak_pairs or peak_pairs <= 0:
            peak_pairs = 148 * 8 * 1.9e9
        # engine 0: popcount scan on the CUDA cores (what the north star names)
        ms0, used0, _ = time_engine(0, 2)
        res0 = (hi.clone(), hd.clone())
        # engine 1 (default at this size): exact E4M3 contraction on the tensor cores
        time_engine(1, 1)
        L.yb_prof_enable(1)
        L.yb_prof_ms(12, None, 1)
        ms1, used1, fb1 = time_engine(1, 3)
        cnt = C.c_long(0)
        ph = {}
        for pid, name in ((12, "expand_codes"), (13, "sample_thresholds"), (14, "e4m3_pass"),
                          (15, "order_and_certify"), (7, "scan_fallback")):
            t = L.yb_prof_ms(pid, C.byref(cnt), 0)
            if cnt.value:
                ph[name] = t / cnt.value
        L.yb_prof_ms(0, None, 1)
        L.yb_prof_enable(0)
        same = bool(torch.equal(res0[0], hi) and torch.equal(res0[1], hd))
        best = min(ms0, ms1) if used1 == 1 else ms0
        pairs = nqh * nbh / (best * 1e-3)
        bf16 = 1637.4
        try:
            bf16 = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("bf16_tflops", bf16)
        except Exception:
            pass
        entry = {
            "queries_per_s": nqh / (best * 1e-3), "ms": best, "pair_distances_per_s": pairs,
            "engines": {
                "popcount_scan": {
                    "ms": ms0, "queries_per_s": nqh / (ms0 * 1e-3),
                    "roofline": {"bound": "popcount pipe: 64-bit xor+popc pair rate MEASURED in this run by a "
                                          "register-only micro-benchmark (yb_debug_popc_pairs_per_s)",
                                 "achieved": nqh * nbh / (ms0 * 1e-3), "peak": peak_pairs,
                                 "frac": nqh * nbh / (ms0 * 1e-3) / peak_pairs, "unit": "pairs/s"}},
            },
            "engines_agree_bit_for_bit": same,
            "note": "algorithmic HBM traffic is 86 MB (13 us at peak): not HBM bound"}
        if used1 == 1:
            kms = ph.get("e4m3_pass", ms1)
            ops = 2.0 * nqh * nbh * 64
            entry["engines"]["tensor_e4m3"] = {
                "ms": ms1, "queries_per_s": nqh / (ms1 * 1e-3), "phase_ms": ph,
                "queries_redone_by_the_scan": int(fb1),
                "roofline": {"bound": "tensor", "kernel": "k_knn_tf32<EPI_LISTS, F8> (kind::f8f6f4, +-1 operands)",
                             "achieved": ops / (kms * 1e-3) / 1e12, "peak": 2.0 * bf16, "unit": "TFLOP/s",
                             "frac": ops / (kms * 1e-3) / 1e12 / (2.0 * bf16),
                             "peak_source": "2 x MEASURED_PEAKS.json bf16_tflops (fp8 dense = twice the bf16 rate)",
                             "algorithmic_flops_per_launch": ops,
                             "note": "K = 64 per tile: two MMAs against a 128 x 256 epilogue, so the pass is "
                                     "epilogue bound by construction; pairs/s vs the popcount ceiling: %.2f"
                                     % (nqh * nbh / (kms * 1e-3) / peak_pairs)}}
        out["hamming_knn_10Mx64bit_10kq_k100"] = entry
        del hb, hq, hi, hd
        torch.cuda.empty_cache()
    except Exception as e:  # extras never break the headline line
        out["hamming_error"] = str(e)
    try:
        from yael_b200.ynumpy import KMEANS_INIT_USER, KMEANS_QUIET
        n, d, k, niter = 10_000_000, 128, 65536, 2
        torch.manual_seed(1237)
        v = torch.rand((n, d), device=dev, dtype=torch.float32)
        cent = v[torch.randperm(n, device=dev)[:k]].cpu().numpy().copy()
        nassign = np.empty(k, np.int32)
        f, i = C.POINTER(C.c_float), C.POINTER(C.c_int)
        L.yb_prof_enable(1)
        L.yb_prof_ms(1, None, 1)
        times = []
        for _ in range(2):
            c = cent.copy()
            torch.cuda.synchronize()
            t = time.perf_counter()
            q = L.yb_kmeans_dev(d, n, k, niter, v.data_ptr(), KMEANS_INIT_USER | KMEANS_QUIET, 1, 1,
                                c.ctypes.data_as(f), None, None, nassign.ctypes.data_as(i), None, None)
            torch.cuda.synchronize()
            times.append((time.perf_counter() - t) / niter)
        cnt = C.c_long(0)
        tms = L.yb_prof_ms(1, C.byref(cnt), 1)
        kms = tms / max(1, cnt.value)
        L.yb_prof_enable(0)
        out["kmeans_10Mx128_k65536"] = {
            "iter_per_s": 1.0 / min(times), "s_per_iter": min(times), "qerr": float(q),
            "assignment_tf32_kernel_ms": kms,
            "assignment_tflops": 2.0 * n * k * d / (kms * 1e-3) / 1e12,
            "note": "one GPU, points resident in HBM, centroids from a seeded permutation (USER init)"}
        del v
        torch.cuda.empty_cache()
    except Exception as e:
        out["kmeans_error"] = str(e)
    return out


def measure_extras_sharded(torch, dist, ydist, L, dev, rank, world):
    """N > 1: the other two BASELINE shapes SHARDED over the ranks, total work fixed:
      - k-means, BASELINE configs[3]: n = 10M points split over the ranks, k = 65536, d = 128; one
        all-reduce of (k*d sums, k counts, qerr) per iteration through the C host loop's hook
      - Hamming kNN, BASELINE configs[2]: 10M codes split over the ranks, all-gather + merge
    Every rank runs the same collectives in the same order; a failure on one rank is turned into a
    flag that all ranks agree on before the next collective."""
    out = {}

    def all_ok(ok):
        t = torch.tensor([1 if ok else 0], device=dev, dtype=torch.int32)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return bool(t.item())

    def max_over_ranks(x):
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- Hamming, database sharded
    try:
        nbh, nqh, kh = 10_000_000, 10_000, 100
        lo, hi = ydist.shard_bounds(nbh, world)[rank]
        g = torch.Generator(device=dev)
        g.manual_seed(1236 + rank)
        hb = torch.randint(0, 256, (hi - lo, 8), device=dev, dtype=torch.uint8, generator=g)
        gq = torch.Generator(device=dev)
        gq.manual_seed(99)
        hq = torch.randint(0, 256, (nqh, 8), device=dev, dtype=torch.uint8, generator=gq)
        sh = ydist.ShardedHamming(hb, kh, rank=rank, world=world, id_offset=lo)
        ok = True
    except Exception as e:
        ok = False
        out["hamming_error"] = str(e)
    if all_ok(ok):
        best = 1e9
        for _ in range(3):
            dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            sh.search(hq)
            e1.record()
            e1.synchronize()
            best = min(best, max_over_ranks(e0.elapsed_time(e1)))
        out["hamming_knn_10Mx64bit_10kq_k100_sharded"] = {
            "queries_per_s": nqh / (best * 1e-3), "ms": best, "ranks": world,
            "engine": int(L.yb_last_hamming_engine()), "scaling": "strong (10M codes split over the ranks)"}
        del hb, hq, sh
        torch.cuda.empty_cache()

    # ---- k-means, points sharded
    try:
        n, d, k, niter = 10_000_000, 128, 65536, 2
        lo, hi = ydist.shard_bounds(n, world)[rank]
        g = torch.Generator(device=dev)
        g.manual_seed(1237 + rank)
        v = torch.rand((hi - lo, d), device=dev, dtype=torch.float32, generator=g)
        cent = v[:k].contiguous() if rank == 0 else torch.empty((k, d), device=dev, dtype=torch.float32)
        ok = v.shape[0] >= k or rank != 0
    except Exception as e:
        ok = False
        out["kmeans_error"] = str(e)
    if all_ok(ok):
        dist.broadcast(cent, 0)
        c0 = cent.cpu().numpy()
        times = []
        okrun = True
        for _ in range(2):
            dist.barrier()
            torch.cuda.synchronize()
            t = time.perf_counter()
            try:
                _, q, _, _ = ydist.sharded_kmeans(v, k, niter, c0, n)
            except Exception as e:  # the all-reduce hook returns an error code instead of raising
                okrun = False
                out["kmeans_error"] = str(e)
            torch.cuda.synchronize()
            times.append(max_over_ranks((time.perf_counter() - t) / niter))
            if not all_ok(okrun):
                break
        if okrun and times:
            out["kmeans_10Mx128_k65536_sharded"] = {
                "iter_per_s": 1.0 / min(times), "s_per_iter": min(times), "ranks": world, "qerr": float(q),
                "scaling": "strong (10M points split over the ranks, one all-reduce per iteration)"}
        del v
        torch.cuda.empty_cache()
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist
    import yael_b200
    from yael_b200 import dist as ydist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    L = yael_b200.lib()
    if L.yb_device_count() <= 0:
        raise SystemExit("bench.py needs a GPU: " + L.yb_last_error().decode())
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    L.yb_set_device(local)
    if world > 1:
        import datetime
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=240))

    # a real (non-default) stream: the library launches on the stream handle it is given, and the
    # CUDA events below must sit on that same stream (handle 0 would mean "the library's own")
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    base_h, query_h = gen_data(rank)
    base = torch.from_numpy(base_h).to(dev)
    query = torch.from_numpy(query_h).to(dev)
    searcher = ydist.ShardedKnn(base, K, rank=rank, world=world)

    def step():
        return searcher.search(query)

    for _ in range(max(3, args.warmup)):
        step()
    torch.cuda.synchronize()
    engine = L.yb_last_knn_engine()
    uncert = L.yb_last_knn_uncertified()

    # ---- timed region: HBM-resident inputs, CUDA events on the launching stream
    L.yb_prof_enable(1)
    L.yb_prof_ms(1, None, 1)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    try:
        uuid = "GPU-" + str(torch.cuda.get_device_properties(dev).uuid)
    except Exception:
        uuid = None
    sampler = ClockSampler(local, uuid)
    sampler.start()
    L.yb_launch_count(1)
    if world > 1:
        searcher.events = []   # per-search CUDA events: local scan / all-gather / merge
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.mark_begin()
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    torch.cuda.synchronize()
    sampler.mark_end()
    if world > 1:
        dist.barrier()
    clocks = sampler.stop()
    launches = L.yb_launch_count(0)
    shard_phase_ms = searcher.phase_ms() if world > 1 else None
    searcher.events = None
    ms = e0.elapsed_time(e1) / args.steps
    cnt = C.c_long(0)
    phase_ms = {}
    for ph, name in ((0, "center_and_norms"), (10, "sample_thresholds"), (1, "tf32_shortlist"),
                     (2, "merge_select"), (3, "rerank"), (4, "exact_fallback"), (5, "exact_slab"),
                     (6, "row_select")):
        t = L.yb_prof_ms(ph, C.byref(cnt), 0)
        if cnt.value:
            phase_ms[name] = t / cnt.value
    L.yb_prof_ms(0, None, 1)
    L.yb_prof_enable(0)
    if world > 1:
        tmax = torch.tensor([ms], device=dev)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        ms = float(tmax.item())
    value = world * NQ / (ms * 1e-3)

    # ---- end to end through the drop-in C call with pinned host buffers
    bh = torch.from_numpy(base_h).pin_memory()
    qh = torch.from_numpy(query_h).pin_memory()
    idx_h = torch.empty((NQ, K), dtype=torch.int32).pin_memory()
    dis_h = torch.empty((NQ, K), dtype=torch.float32).pin_memory()
    e2e_steps = max(2, min(args.steps, 5))

    def e2e_step():
        searcher.search_host(bh.numpy(), qh.numpy(), idx_h.numpy(), dis_h.numpy())

    e2e_step()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    t_e2e = (time.perf_counter() - t0) / e2e_steps
    if world > 1:
        tmax = torch.tensor([t_e2e], device=dev)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        t_e2e = float(tmax.item())
    e2e = {"value": world * NQ / t_e2e, "unit": UNIT,
           "h2d_bytes_per_step": int(base_h.nbytes + query_h.nbytes),
           "d2h_bytes_per_step": int(NQ * K * 8), "ms_per_step": t_e2e * 1e3}

    extras_sharded = None
    if world > 1 and not args.no_extras:
        try:
            extras_sharded = measure_extras_sharded(torch, dist, ydist, L, dev, rank, world)
        except Exception as e:
            extras_sharded = {"error": str(e)}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    # denominator: MEASURED_PEAKS.json holds the dense bf16 GEMM rate of this pool's B200s ("of
    # measured").  The resident passes read FP16 operands (kind::f16 runs at the bf16 rate); if the
    # library fell back to TF32 operands (kind::tf32, half that rate) the peak is bf16_tflops / 2.
    # The cuBLAS TF32 GEMM rate measured in this very run is reported beside it.
    tf32_meas = measure_tf32_peak(torch, dev) if world == 1 else None
    bf16_peak = peaks.get("bf16_tflops", 1590.0)
    operands = "fp16" if L.yb_last_knn_operands() == 2 else "tf32"
    peak = bf16_peak if operands == "fp16" else bf16_peak / 2.0
    src = "MEASURED_PEAKS.json bf16_tflops" if peaks else "fallback 1590 TF/s bf16"
    peak_src = (src + (" (FP16 operands run at the bf16 rate)" if operands == "fp16" else
                       " / 2 (TF32 dense = half the bf16 rate)") + (", of measured" if peaks else ", of fallback"))
    roof = None
    if "tf32_shortlist" in phase_ms:
        kms = phase_ms["tf32_shortlist"]
        ach = 2.0 * NQ * NB * D / (kms * 1e-3) / 1e12
        traffic = None
        try:  # DRAM bytes of this kernel from the committed ncu --set full capture
            traffic = json.load(open(os.path.join(ROOT, "profiles", "r1_knn_tf32_traffic.json")))["dram_bytes_per_launch"]
        except Exception:
            pass
        roof = {"bound": "tensor",
                "kernel": "k_knn_tf32<EPI_LISTS, %s> (full pass; the sampling passes are in phase_ms)" % operands,
                "operands": operands,
                "note": ("FP16 operands (half as many MMAs per tile as TF32) with |b|^2 folded into the "
                         "contraction: the epilogue is a MAX tree over raw accumulators; ncu: tensor pipe 52 % "
                         "busy; the pass is paced by the fused top-k epilogue and the operand feed, not by the MMA "
                         "rate -- decomposition in DESIGN.md 5.1 and profiles/README.md"),
                "achieved": ach, "peak": peak,
                "unit": "TFLOP/s", "frac": ach / peak, "traffic": traffic,
                "kernel_ms": kms, "peak_source": peak_src,
                "frac_of_bf16_peak": ach / bf16_peak,
                "cublas_tf32_tflops_this_run": tf32_meas,
                "frac_of_cublas_tf32": (ach / tf32_meas) if tf32_meas else None,
                "algorithmic_flops_per_launch": 2.0 * NQ * NB * D}
    elif "exact_slab" in phase_ms:
        kms = phase_ms["exact_slab"]
        roof = {"bound": "tensor", "kernel": "k_l2_simt (exact FP32 engine, CUDA cores)",
                "achieved": 2.0 * NQ * NB * D / (kms * 1e-3) / 1e12, "peak": peak, "unit": "TFLOP/s",
                "frac": 2.0 * NQ * NB * D / (kms * 1e-3) / 1e12 / peak, "traffic": None,
                "peak_source": peak_src}

    cpu = cpu_reference_rate(base_h, query_h)

    extras = extras_sharded
    if world == 1 and not args.no_extras:
        extras = measure_extras(torch, L, dev)

    print(json.dumps({
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(3, args.warmup), "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None,
        "dtype": "f32 (%s tensor-core shortlist with FP32 accumulation, exact f32 re-rank)" % operands,
        "data": "synthetic uniform[0,1), seeds 1234/1235",
        "config": {
            "workload": WORKLOAD,
            "parallelism": ("single GPU" if world == 1 else
                            "database sharded x%d (1M rows per rank), NCCL all-gather of per-rank "
                            "top-k + merge; value counts query x 1M-shard scans" % world),
            "l2": "database (512 MB) is larger than L2 (126 MB): no flush needed between steps",
            "engine": ("tcgen05 %s + FP32 re-rank" % operands) if engine == 1 else "exact FP32 SIMT",
            "uncertified_queries_redone_exactly": int(uncert),
            "phase_ms": phase_ms,
            "shard_phase_ms_rank0": shard_phase_ms,
        },
        "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
        "roofline": roof, "cpu_baseline": cpu, "extras": extras,
    }))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-extras", action="store_true",
                    help="skip the k-means / Hamming side measurements (N=1 only)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
