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

    def __init__(self, gpu_index, uuid=None, poll_ms=2.0, enabled=True):
        # N > 1: NVML calls from 8 processes at 500 Hz each contend inside the driver (measured at
        # N = 8: a poll iteration took ~15 ms instead of 2 and the timed loop ran at 9.2 ms per step
        # against 4.1 ms without the sampler), so only rank 0 samples there, at a lower rate
        self.poll_s = poll_ms * 1e-3
        self.enabled = enabled
        self.gpu = gpu_index
        self.uuid = uuid
        self.samples = []   # (t, sm_mhz, sm_max_mhz, reason_bits)
        self.lines = []
        self.proc = None
        self.nvml = None
        self.stop_flag = False
        self.t0 = self.t1 = None

    def start(self):
        if not self.enabled:
            return
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
            time.sleep(self.poll_s)

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def mark_begin(self):
        self.t0 = time.perf_counter()

    def mark_end(self):
        self.t1 = time.perf_counter()

    def stop(self):
        if not self.enabled:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["not sampled on this rank"]}
        if self.nvml is not None:
            self.stop_flag = True
            self.th.join(timeout=1)
            sel = [x for x in self.samples
                   if (self.t0 is None or x[0] >= self.t0) and (self.t1 is None or x[0] <= self.t1)]
            if not sel:
                return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
            reasons = sorted({name for name, bit in self.REASONS for x in sel if x[3] & bit})
            return {"sm_mhz": float(np.median([x[1] for x in sel])), "sm_max_mhz": float(sel[0][2]),
                    "reasons": reasons, "samples": len(sel), "source": "nvml, %g ms polling" % (self.poll_s * 1e3)}
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
    bounded sample: all host threads, OPENBLAS_NUM_THREADS=1 (yael threads itself).  Returns the
    cpu_baseline object and the reference's ANSWER for the sampled queries (n, idx, dis), which
    run_ours compares with the GPU result outside any timed region."""
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
    widx, wdis = fn(base, query[:n])
    dt = time.perf_counter() - t
    return ({"value": n / dt, "unit": UNIT, "cores": cores, "kind": kind,
             "sample": "%d of %d queries against the full 1M x 128 database, k=100, %.1f s" % (n, NQ, dt)},
            (n, widx, wdis))


def knn_parity(idx, dis, widx, wdis, base, query, config):
    """GPU result against the reference's for the same queries (north-star tolerance: distances
    within 1e-5 relative, ids identical except for ties inside that tolerance).  A differing id is
    accepted only if its OWN distance to the query, recomputed in float64, is within the tolerance
    of the reference's distance at that rank."""
    n = widx.shape[0]
    idx, dis = idx[:n], dis[:n]
    valid = widx >= 0
    rel = np.abs(dis[valid].astype(np.float64) - wdis[valid]) / np.maximum(np.abs(wdis[valid]), 1e-30)
    diff = (idx != widx) & valid
    ties_ok = True
    b64 = None
    for qi, j in zip(*np.nonzero(diff)):
        own = ((base[idx[qi, j]].astype(np.float64) - query[qi].astype(np.float64)) ** 2).sum()
        if abs(own - float(wdis[qi, j])) > 1e-5 * max(abs(float(wdis[qi, j])), 1e-6) + 2e-6 * float(
                (query[qi].astype(np.float64) ** 2).sum()):
            ties_ok = False
    dup_free = all(len(set(r.tolist())) == len(r) for r in idx[np.unique(np.nonzero(diff)[0])])
    del b64
    return {"config": config, "n_queries": int(n), "checked_against": "reference CPU result of cpu_baseline",
            "ids_identical": int((~diff).all(axis=1).sum()), "ids_differing_entries": int(diff.sum()),
            "ids_equal_outside_ties": bool(ties_ok and dup_free and np.array_equal(valid, idx >= 0)),
            "max_rel_dis": float(rel.max()) if rel.size else 0.0,
            "distances_bit_identical": bool(np.array_equal(dis[valid], wdis[valid])),
            "ok": bool(ties_ok and dup_free and (rel.max() if rel.size else 0.0) <= 1e-5)}


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
def _bind_to_gpu_numa_node(torch, dev):
    """Restrict this process to the CPUs of the NUMA node the GPU is attached to (sysfs), so that the
    pinned host buffers allocated next are node-local.  Returns the node, or None when the topology is
    not exposed / the node's CPUs are not available to this process."""
    try:
        bus = torch.cuda.get_device_properties(dev).pci_bus_id
        dom = getattr(torch.cuda.get_device_properties(dev), "pci_domain_id", 0)
        devid = getattr(torch.cuda.get_device_properties(dev), "pci_device_id", 0)
        path = "/sys/bus/pci/devices/%04x:%02x:%02x.0/numa_node" % (dom, bus, devid)
        node = int(open(path).read().strip())
        if node < 0:
            return None
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return node
    except Exception:
        return None


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


def _peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


def _phase_means(L, names):
    cnt = C.c_long(0)
    out = {}
    for pid, name in names:
        t = L.yb_prof_ms(pid, C.byref(cnt), 0)
        if cnt.value:
            out[name] = t / cnt.value
    return out


def measure_hamming(torch, L, dev):
    """BASELINE configs[2]: Hamming k-NN, 10M x 64-bit packed codes, 10k queries, k = 100, one GPU.
    value: yb_nn_hamming on HBM-resident codes (CUDA events).  e2e: the drop-in nn_hamming() on
    pinned HOST buffers (copies inside the timed region).  roofline: both engines.  cpu_baseline:
    the reference's compute_hamming (yael/hamming.c:177-219) over database blocks + stable select
    on a >= 1e9-pair slice, all host threads.  parity: the GPU answer for that slice's queries
    against the CPU answer, bit for bit."""
    nbh, nqh, kh = 10_000_000, 10_000, 100
    torch.manual_seed(1236)
    hb = torch.randint(0, 256, (nbh, 8), device=dev, dtype=torch.uint8)
    hq = torch.randint(0, 256, (nqh, 8), device=dev, dtype=torch.uint8)
    hi = torch.empty((nqh, kh), device=dev, dtype=torch.int32)
    hd = torch.empty((nqh, kh), device=dev, dtype=torch.int16)
    sp = C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def time_engine(engine, reps):
        L.yb_set_hamming_engine(engine)
        best = 1e9
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            rc = L.yb_nn_hamming(nqh, nbh, 8, kh, hb.data_ptr(), hq.data_ptr(), hi.data_ptr(), hd.data_ptr(), 0, sp)
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
    time_engine(0, 1)                      # warm-up
    L.yb_prof_enable(1)
    L.yb_prof_ms(7, None, 1)
    ms0, used0, _ = time_engine(0, 3)      # engine 0: popcount scan on the CUDA cores
    ph0 = _phase_means(L, ((7, "popcount_scan"),))
    L.yb_prof_ms(0, None, 1)
    res0 = (hi.clone(), hd.clone())
    for _ in range(3):
        time_engine(1, 1)                  # engine 1 warm-up (>= 3 steps)
    L.yb_prof_ms(12, None, 1)
    ms1, used1, fb1 = time_engine(1, 5)    # engine 1: exact E4M3 contraction on the tensor cores
    ph = _phase_means(L, ((12, "expand_codes"), (13, "sample_thresholds"), (14, "e4m3_pass"),
                          (15, "order_and_certify"), (7, "scan_fallback")))
    L.yb_prof_ms(0, None, 1)
    L.yb_prof_enable(0)
    gpu_idx = hi.cpu().numpy()
    gpu_dis = hd.cpu().numpy().view(np.uint16)
    same = bool(torch.equal(res0[0], hi) and torch.equal(res0[1], hd))
    best = min(ms0, ms1) if used1 == 1 else ms0
    bf16 = _peaks().get("bf16_tflops", 1590.0)
    scan_ms = ph0.get("popcount_scan", ms0)
    block = {
        "metric": "Hamming kNN queries/s (10M x 64-bit codes, 10k queries, k=100)",
        "value": nqh / (best * 1e-3), "unit": "queries/s", "ms_per_step": best,
        "dtype": "u8 codes, u16 distances (bit-exact)",
        "config": {"workload": "Hamming kNN: 10M x 64-bit packed codes, 10000 queries, k=100 (BASELINE configs[2])",
                   "engine_used": "tensor_e4m3" if (used1 == 1 and ms1 <= ms0) else "popcount_scan",
                   "l2": "80 MB of codes fit in L2; expanded operands (640 MB) do not"},
        "pair_distances_per_s": nqh * nbh / (best * 1e-3),
        "engines_agree_bit_for_bit": same,
        "roofline": {"bound": "hbm", "kernel": "k_nn_hamming_scan (engine 0, what the north star names)",
                     "achieved": 86e6 / (scan_ms * 1e-3) / 1e9, "peak": _peaks().get("hbm_gbs", 6450.0),
                     "unit": "GB/s", "frac": 86e6 / (scan_ms * 1e-3) / 1e9 / _peaks().get("hbm_gbs", 6450.0),
                     "traffic": None, "kernel_ms": scan_ms,
                     "algorithmic_bytes_per_launch": 86e6,
                     "note": "HBM is NOT what bounds this scan (86 MB algorithmic = 13 us at peak): at 10k "
                             "queries it is bound by the popcount pipe -- see roofline_popc"},
        "roofline_popc": {"bound": "popcount pipe: 64-bit xor+popc pair rate MEASURED in this run by a "
                                   "register-only micro-benchmark (yb_debug_popc_pairs_per_s)",
                          "kernel": "k_nn_hamming_scan", "achieved": nqh * nbh / (scan_ms * 1e-3),
                          "peak": peak_pairs, "frac": nqh * nbh / (scan_ms * 1e-3) / peak_pairs,
                          "unit": "pairs/s", "kernel_ms": scan_ms},
        "engines": {"popcount_scan": {"ms": ms0, "queries_per_s": nqh / (ms0 * 1e-3)}},
    }
    if used1 == 1:
        kms = ph.get("e4m3_pass", ms1)
        ops = 2.0 * nqh * nbh * 64
        block["engines"]["tensor_e4m3"] = {"ms": ms1, "queries_per_s": nqh / (ms1 * 1e-3), "phase_ms": ph,
                                           "queries_redone_by_the_scan": int(fb1)}
        block["roofline_tensor"] = {
            "bound": "tensor", "kernel": "k_knn_tf32<EPI_HAMP, F8> (kind::f8f6f4, +-1 operands)",
            "achieved": ops / (kms * 1e-3) / 1e12, "peak": 2.0 * bf16, "unit": "TFLOP/s",
            "frac": ops / (kms * 1e-3) / 1e12 / (2.0 * bf16), "kernel_ms": kms,
            "peak_source": "2 x MEASURED_PEAKS.json bf16_tflops (fp8 dense = twice the bf16 rate)",
            "algorithmic_flops_per_launch": ops,
            "pairs_per_s_vs_popcount_ceiling": nqh * nbh / (kms * 1e-3) / peak_pairs}

    # ---- e2e: nn_hamming() on pinned host buffers
    hb_h = hb.cpu().pin_memory()
    hq_h = hq.cpu().pin_memory()
    oi_h = torch.empty((nqh, kh), dtype=torch.int32).pin_memory()
    od_h = torch.empty((nqh, kh), dtype=torch.int16).pin_memory()
    u8, u16, i32 = C.POINTER(C.c_uint8), C.POINTER(C.c_uint16), C.POINTER(C.c_int)

    def e2e_step():
        L.nn_hamming(nqh, nbh, 8, kh, C.cast(hb_h.data_ptr(), u8), C.cast(hq_h.data_ptr(), u8),
                     C.cast(oi_h.data_ptr(), i32), C.cast(od_h.data_ptr(), u16))

    for _ in range(3):
        e2e_step()
    t0 = time.perf_counter()
    for _ in range(5):
        e2e_step()
    te = (time.perf_counter() - t0) / 5
    block["e2e"] = {"value": nqh / te, "unit": "queries/s", "ms_per_step": te * 1e3,
                    "h2d_bytes_per_step": int(nbh * 8 + nqh * 8), "d2h_bytes_per_step": int(nqh * kh * 6),
                    "call": "nn_hamming() on pinned host buffers"}
    e2e_same = bool(np.array_equal(oi_h.numpy(), gpu_idx) and np.array_equal(od_h.numpy().view(np.uint16), gpu_dis))

    # ---- CPU baseline + parity on a >= 1e9-pair slice (100 queries x 10M codes)
    try:
        from oracle import bindings as ob
        ns = 100
        cores = len(os.sched_getaffinity(0))
        t = time.perf_counter()
        widx, wdis, kind = ob.ref_nn_hamming_blocked(hb_h.numpy(), hq_h.numpy()[:ns], kh, threads=cores)
        dt = time.perf_counter() - t
        block["cpu_baseline"] = {"value": ns / dt, "unit": "queries/s", "cores": cores, "kind": kind,
                                 "sample": "%d of %d queries x 10M codes (%.1e pairs): compute_hamming over 1M-code "
                                           "blocks on %d threads + stable (distance, id) select, %.1f s"
                                           % (ns, nqh, ns * nbh, cores, dt)}
        block["parity"] = {"config": "C3: 10M x 64-bit codes, k=100", "n_queries": ns,
                           "pairs_checked": int(ns * nbh),
                           "ids_bit_identical": bool(np.array_equal(gpu_idx[:ns], widx)),
                           "distances_bit_identical": bool(np.array_equal(gpu_dis[:ns], wdis)),
                           "e2e_call_equals_resident_call": e2e_same,
                           "ok": bool(np.array_equal(gpu_idx[:ns], widx) and np.array_equal(gpu_dis[:ns], wdis) and same)}
    except Exception as e:
        block["cpu_baseline_error"] = str(e)
    del hb, hq, hi, hd, hb_h, hq_h
    torch.cuda.empty_cache()
    return block


def measure_kmeans(torch, L, dev, niter=10):
    """BASELINE configs[3] on one GPU: k-means, n = 10M, d = 128, k = 65536, `niter` iterations from a
    seeded permutation of the points (KMEANS_INIT_USER).  value: iterations/s with the points
    resident in HBM (yb_kmeans_dev).  e2e: the drop-in kmeans() on pinned HOST points.  roofline:
    the assignment's tensor pass (tensor) and the centroid update (HBM).  cpu_baseline: the
    reference's assignment (knn_full_thread k = 1, yael/kmeans.c:242) on a slice of the points
    against all 65536 centroids, scaled to an iteration; plus the reference's kmeans() on BASELINE
    configs[0].  parity: the teacher-forced GPU step on that slice against the reference's."""
    from yael_b200.ynumpy import KMEANS_INIT_USER, KMEANS_QUIET
    n, d, k = 10_000_000, 128, 65536
    torch.manual_seed(1237)
    v = torch.rand((n, d), device=dev, dtype=torch.float32)
    cent0 = v[torch.randperm(n, device=dev)[:k]].cpu().numpy().copy()
    nassign = np.empty(k, np.int32)
    f, i = C.POINTER(C.c_float), C.POINTER(C.c_int)
    flags = KMEANS_INIT_USER | KMEANS_QUIET

    def run(niter_, ptr):
        c = cent0.copy()
        torch.cuda.synchronize()
        t = time.perf_counter()
        q = L.yb_kmeans_dev(d, n, k, niter_, ptr, flags, 1, 1, c.ctypes.data_as(f), None, None,
                            nassign.ctypes.data_as(i), None, None)
        torch.cuda.synchronize()
        return time.perf_counter() - t, q, c

    run(3, v.data_ptr())                                     # warm-up: 3 iterations
    L.yb_prof_enable(1)
    L.yb_prof_ms(1, None, 1)
    L.yb_launch_count(1)
    # SM clocks during the timed iterations: 1.6 s of back-to-back tensor passes run at the
    # power-capped clock, which is what MEASURED_PEAKS.json's SUSTAINED bf16 figure was taken at
    try:
        uuid = "GPU-" + str(torch.cuda.get_device_properties(dev).uuid)
    except Exception:
        uuid = None
    sampler = ClockSampler(dev.index or 0, uuid, poll_ms=20.0,
                           enabled=os.environ.get("BENCH_CLOCKS", "on") != "off")
    sampler.start()
    sampler.mark_begin()
    tt, q, cent = run(niter, v.data_ptr())
    sampler.mark_end()
    km_clocks = sampler.stop()
    launches = L.yb_launch_count(0)
    ph = _phase_means(L, ((0, "center_and_convert"), (1, "tensor_pass"), (3, "rerank_k1"), (4, "exact_fallback"),
                          (8, "update_sort_and_sums"), (9, "scale"), (18, "update_row_stream_k_segsum_sorted")))
    L.yb_prof_ms(0, None, 1)
    L.yb_prof_enable(0)
    s_iter = tt / niter
    pk = _peaks()
    bf16, hbm = pk.get("bf16_tflops", 1590.0), pk.get("hbm_gbs", 6450.0)
    operands = "fp16" if L.yb_last_knn_operands() == 2 else "tf32"
    # the pass runs back to back for seconds (10 x 150 ms): the SUSTAINED measured rate is its
    # denominator (B200_PROFILING.md: burst for a kernel timed alone, sustained inside a long step);
    # the fraction of the burst rate is reported beside it
    bf16_s = pk.get("bf16_tflops_sustained", bf16)
    tpeak = bf16_s if operands == "fp16" else bf16_s / 2.0
    tburst = bf16 if operands == "fp16" else bf16 / 2.0
    kms = ph.get("tensor_pass", s_iter * 1e3)
    flops = 2.0 * n * k * d
    ubytes = 4.0 * n * d + 4.0 * n + 4.0 * k * d + 4.0 * k
    ums = ph.get("update_sort_and_sums")
    seg_ms = ph.get("update_row_stream_k_segsum_sorted")
    block = {
        "metric": "k-means iter/s (10M x 128, k=65536)", "value": 1.0 / s_iter, "unit": "iter/s",
        "ms_per_step": s_iter * 1e3, "steps": niter, "warmup": 3, "qerr": float(q),
        "dtype": "f32 (%s tensor-core shortlist, exact f32 re-rank and sums)" % operands,
        "config": {"workload": "k-means at scale: n=10M, d=128, k=65536, %d iterations (BASELINE configs[3]), one GPU" % niter,
                   "init": "KMEANS_INIT_USER: rows of a seeded permutation", "phase_ms": ph,
                   "uncertified_points_last_iteration": int(L.yb_last_knn_uncertified())},
        "gpu_launches": int(launches),
        "roofline": {"bound": "tensor", "kernel": "k_knn_tf32<EPI_NEAREST, %s> (assignment, k = 1 margin mode)" % operands,
                     "achieved": flops / (kms * 1e-3) / 1e12, "peak": tpeak, "unit": "TFLOP/s",
                     "frac": flops / (kms * 1e-3) / 1e12 / tpeak, "traffic": None, "kernel_ms": kms,
                     "algorithmic_flops_per_launch": flops,
                     "frac_of_burst_peak": flops / (kms * 1e-3) / 1e12 / tburst, "burst_peak": tburst,
                     "peak_source": ("MEASURED_PEAKS.json bf16_tflops_sustained (the pass runs back to back for "
                                     "%.1f s; clocks below)" % tt) if pk else "fallback 1590 TF/s bf16"},
        "clocks": km_clocks,
    }
    if ums:
        # the dominant kernel (k_segsum_sorted: the one pass over the points) on its own event-timed
        # duration; the whole update phase (histogram, id scatter, the row stream, qerr) beside it
        kms_u = seg_ms if seg_ms else ums
        block["roofline_update"] = {
            "bound": "hbm",
            "kernel": ("k_segsum_sorted (per-segment id sort + row sums in the reference's order)" if seg_ms
                       else "centroid update: histogram + id scatter + k_segsum_sorted + qerr"),
            "achieved": ubytes / (kms_u * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s",
            "frac": ubytes / (kms_u * 1e-3) / 1e9 / hbm, "traffic": None, "kernel_ms": kms_u,
            "algorithmic_bytes_per_launch": ubytes,
            "whole_update_phase_ms": ums, "whole_update_phase_frac": ubytes / (ums * 1e-3) / 1e9 / hbm,
            "peak_source": "MEASURED_PEAKS.json hbm_gbs" if pk else "fallback 6450 GB/s"}

    # ---- parity + CPU baseline: a slice of the points, teacher-forced against the same centroids
    try:
        from oracle import bindings as ob
        cores = len(os.sched_getaffinity(0))
        ns = 100_000
        vs = v[:ns].contiguous()
        ga = torch.empty(ns, device=dev, dtype=torch.int32)
        gd = torch.empty(ns, device=dev, dtype=torch.float32)
        cd = torch.from_numpy(cent0).to(dev)
        sp = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        rc = L.yb_knn_l2(ns, k, d, 1, cd.data_ptr(), vs.data_ptr(), None, ga.data_ptr(), gd.data_ptr(), 0, sp)
        assert rc == 0, L.yb_last_error()
        sums = torch.empty((k, d), device=dev, dtype=torch.float32)
        cnts = torch.empty(k, device=dev, dtype=torch.int32)
        qe = torch.empty(1, device=dev, dtype=torch.float64)
        rc = L.yb_kmeans_accumulate(d, ns, k, vs.data_ptr(), ga.data_ptr(), gd.data_ptr(), sums.data_ptr(),
                                    cnts.data_ptr(), qe.data_ptr(), 0, sp)
        assert rc == 0, L.yb_last_error()
        torch.cuda.synchronize()
        vs_h = vs.cpu().numpy()
        t = time.perf_counter()
        if ob.have_ref():
            wa, wd = ob.ref_knn(cent0, vs_h, 1, nt=cores)
            kind = "reference"
        else:
            wa, wd = ob.orc_knn(cent0, vs_h, 1, nt=cores)
            kind = "port"
        dt = time.perf_counter() - t
        wa, wd = wa[:, 0], wd[:, 0]
        a, dd = ga.cpu().numpy(), gd.cpu().numpy()
        diff = a != wa
        own_ok = True
        for j in np.nonzero(diff)[0]:
            own = ((cent0[a[j]].astype(np.float64) - vs_h[j].astype(np.float64)) ** 2).sum()
            if abs(own - float(wd[j])) > 1e-5 * abs(float(wd[j])) + 2e-6 * float((vs_h[j].astype(np.float64) ** 2).sum()):
                own_ok = False
        rel = np.abs(dd.astype(np.float64) - wd) / np.maximum(np.abs(wd), 1e-30)
        # centroid sums of the slice under the REFERENCE's assignment, in float64
        ws = np.zeros((k, d), np.float64)
        np.add.at(ws, wa, vs_h.astype(np.float64))
        wc = np.bincount(wa, minlength=k)
        gs, gc = sums.cpu().numpy(), cnts.cpu().numpy()
        same_rows = (gc == wc) & (gc > 0)
        if diff.any():  # rows touched by a flipped tie differ legitimately
            touched = np.zeros(k, bool)
            touched[a[diff]] = True
            touched[wa[diff]] = True
            same_rows &= ~touched
        cerr = np.abs(gs[same_rows] / gc[same_rows, None] - ws[same_rows] / wc[same_rows, None]).max() if same_rows.any() else 0.0
        block["parity"] = {"config": "C4: teacher-forced step, %d-point slice x 65536 centroids, d=128" % ns,
                           "n_points": ns, "checked_against": "%s knn_full_thread(k=1) + float64 sums" % kind,
                           "assign_identical": int((~diff).sum()), "assign_differing": int(diff.sum()),
                           "assign_equal_outside_ties": bool(own_ok), "max_rel_dis": float(rel.max()),
                           "distances_bit_identical": bool(np.array_equal(dd, wd)),
                           "max_abs_centroid_err_vs_float64": float(cerr),
                           "ok": bool(own_ok and rel.max() <= 1e-5 and cerr <= 1e-4)}
        cpu_iter_s = dt * (n / ns)
        block["cpu_baseline"] = {"value": 1.0 / cpu_iter_s, "unit": "iter/s", "cores": cores, "kind": kind,
                                 "sample": "assignment (knn_full_thread, k=1: >99%% of an iteration) of %d of the 10M points "
                                           "against all 65536 centroids on %d threads: %.1f s, scaled x%d to one iteration"
                                           % (ns, cores, dt, n // ns)}
        # BASELINE configs[0], the reference's own CPU-runnable case, in full
        v1 = np.random.RandomState(1234).random_sample((100000, 128)).astype(np.float32)
        t = time.perf_counter()
        if ob.have_ref():
            ob.ref_kmeans(v1, 256, 20, ob.KMEANS_QUIET | cores, 1234)
        else:
            ob.orc_kmeans(v1, 256, 20, ob.KMEANS_QUIET | cores, 1234)
        dt1 = time.perf_counter() - t
        v1d = torch.from_numpy(v1).to(dev)
        c1 = np.zeros((256, 128), np.float32)
        torch.cuda.synchronize()
        best1 = 1e9
        for _ in range(3):
            t = time.perf_counter()
            L.yb_kmeans_dev(128, 100000, 256, 20, v1d.data_ptr(), KMEANS_QUIET, 1234, 1, c1.ctypes.data_as(f),
                            None, None, None, None, None)
            torch.cuda.synchronize()
            best1 = min(best1, time.perf_counter() - t)
        block["config0_kmeans_100k_k256_20it"] = {"cpu_iter_per_s": 20.0 / dt1, "cpu_kind": kind, "cpu_cores": cores,
                                                  "gpu_iter_per_s": 20.0 / best1}
        del vs, ga, gd, cd, sums, cnts, v1d
    except Exception as e:
        block["cpu_baseline_error"] = str(e)

    # ---- e2e: the drop-in kmeans() on pinned host points
    try:
        vh = torch.empty((n, d), dtype=torch.float32).pin_memory()
        vh.copy_(v)
        torch.cuda.synchronize()

        def e2e_call(niter_):
            c = cent0.copy()
            t = time.perf_counter()
            L.kmeans(d, n, k, niter_, C.cast(vh.data_ptr(), f), flags, 1, 1, c.ctypes.data_as(f), None, None,
                     nassign.ctypes.data_as(i))
            return time.perf_counter() - t, c

        e2e_call(1)
        te, ce = e2e_call(niter)
        block["e2e"] = {"value": niter / te, "unit": "iter/s", "ms_per_step": te / niter * 1e3,
                        "h2d_bytes_per_step": int((n * d * 4 + k * d * 4) / niter),
                        "d2h_bytes_per_step": int((k * d * 4 + k * 4) / niter),
                        "call": "kmeans() on pinned host points, %d iterations per call: the 5.12 GB of points "
                                "cross PCIe once per call" % niter,
                        "equals_resident_run": bool(np.array_equal(ce, cent))}
        del vh
    except Exception as e:
        block["e2e_error"] = str(e)
    del v
    torch.cuda.empty_cache()
    return block


def measure_cross(torch, L, dev):
    """compute_cross_distances (yael/nn.c:92-129) at 10 000 x 100 000 x 128: the tensor-core engine
    (split-precision FP16 operands, both norms folded into the contraction; output-bound: 4 B per pair)
    beside the exact FP32 engine, CUDA events on the launching stream, and a parity slice against the
    oracle (north star: within 1e-5 relative)."""
    from oracle import bindings as ob
    na, nb, d = 10_000, 100_000, 128
    g = torch.Generator(device=dev)
    g.manual_seed(4242)
    a = torch.rand((na, d), device=dev, generator=g)
    b = torch.rand((nb, d), device=dev, generator=g)
    out = torch.empty((nb, na), device=dev)
    st = torch.cuda.current_stream()
    sp = C.c_void_p(st.cuda_stream)
    res = {}
    kernel_ms = None
    for engine, name in ((1, "tensor_split_fp16"), (0, "exact_fp32_simt")):
        L.yb_set_cross_engine(engine)
        for _ in range(2):
            rc = L.yb_cross_distances_l2(d, na, nb, a.data_ptr(), d, b.data_ptr(), d, out.data_ptr(), na, sp)
            assert rc == 0, L.yb_last_error()
        torch.cuda.synchronize()
        reps = 5
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(reps):
            L.yb_cross_distances_l2(d, na, nb, a.data_ptr(), d, b.data_ptr(), d, out.data_ptr(), na, sp)
        e1.record(st)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        res[name] = {"ms": ms, "engine_used": int(L.yb_last_cross_engine()),
                     "pairs_per_s": na * nb / (ms * 1e-3), "output_GBps": 4.0 * na * nb / (ms * 1e-3) / 1e9}
        if engine == 1:
            got = out[:2000, :512].cpu().numpy()
            # the contraction kernel alone (yb_prof phase 1) and the operand preparation (phase 0),
            # CUDA events on the launching stream, in a separate profiled round
            cnt = C.c_long(0)
            L.yb_prof_enable(1)
            for pid in (0, 1):
                L.yb_prof_ms(pid, C.byref(cnt), 1)
            for _ in range(3):
                L.yb_cross_distances_l2(d, na, nb, a.data_ptr(), d, b.data_ptr(), d, out.data_ptr(), na, sp)
            torch.cuda.synchronize()
            ph = _phase_means(L, ((0, "split_and_scale_operands"), (1, "k_knn_2sm_cross")))
            L.yb_prof_enable(0)
            res[name]["phase_ms"] = ph
            kernel_ms = ph.get("k_knn_2sm_cross")
    L.yb_set_cross_engine(-1)
    ah, bh = a[:512].cpu().numpy(), b[:2000].cpu().numpy()
    want = ob.orc_cross(ah, bh, ob.DOT_F32_SEQ)
    rel = np.abs(got - want) / np.maximum(np.abs(want), 1e-30)
    peaks = _peaks()
    hbm = peaks.get("hbm_gbs", 6450.0)
    t = res["tensor_split_fp16"]
    kms = kernel_ms if kernel_ms else t["ms"]
    kgbs = (4.0 * na * nb + 2.0 * (3 * d + 16) * (na + nb)) / (kms * 1e-3) / 1e9
    return {"metric": "compute_cross_distances pairs/s (10k x 100k x 128)", "value": t["pairs_per_s"], "unit": "pairs/s",
            "ms_per_step": t["ms"], "engines": res,
            "roofline": {"bound": "hbm",
                         "kernel": "k_knn_2sm<EPI_CROSS> (output: 4 B per pair through staged TMA stores; split-FP16 "
                                   "operands 2 (3 d + 16) B per row)",
                         "achieved": kgbs, "peak": hbm, "unit": "GB/s", "frac": kgbs / hbm,
                         "kernel_ms": kms, "whole_call_GBps": t["output_GBps"],
                         "traffic": None,
                         "algorithmic_bytes_per_launch": 4.0 * na * nb + 2.0 * (3 * d + 16) * (na + nb),
                         "note": "peak = the measured COPY bandwidth (MEASURED_PEAKS.json); a pure-write kernel "
                                 "(torch fill_) reaches 7.5 TB/s on these boxes (scripts/prof_write_bw.py)"},
            "parity": {"config": "512 x 2000 slice of the 10k x 100k matrix vs the oracle (FP32 FMA chain)",
                       "max_rel_dis": float(rel.max()), "ok": bool(rel.max() <= 1e-5)}}


def measure_consumers(torch, L, dev):
    """The consumers of SURVEY.md 8(f)-N4 that sit on the path's kernels, timed through the drop-in C
    API on device-resident inputs and checked against the oracle on a slice: hkm_quantize (3 levels of
    exact k = 1 searches among 10 children, yael/hkm.c:144-162) and the GMM E-step gmm_compute_p
    (yael/gmm.c:211-367: two contractions + softmax)."""
    from oracle import bindings as ob
    from yael_b200 import _lib as yl
    f = C.POINTER(C.c_float)
    r = np.random.RandomState(77)
    out = {}
    # hkm_quantize: 1M x 128 points, bf = 10, 3 levels
    n, d, bf, nl = 1_000_000, 128, 10, 3
    levels = [r.rand(bf ** (l + 1), d).astype(np.float32) for l in range(nl)]
    ptrs = (f * nl)(*[x.ctypes.data_as(f) for x in levels])
    h = yl.HkmT(nl, bf, bf ** nl, d, C.cast(ptrs, C.POINTER(f)))
    v = torch.rand((n, d), device=dev)
    idx = torch.empty(n, dtype=torch.int32, device=dev)
    vp, ip = C.cast(v.data_ptr(), f), C.cast(idx.data_ptr(), C.POINTER(C.c_int))
    L.hkm_quantize(C.byref(h), n, vp, ip)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        L.hkm_quantize(C.byref(h), n, vp, ip)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / 3 * 1e3
    sl = 20000
    want = ob.orc_hkm_quantize(levels, bf, v[:sl].cpu().numpy())
    out["hkm_quantize"] = {"config": "1M x 128 points, bf = 10, 3 levels (device-resident points)", "ms": ms,
                           "points_per_s": n / (ms * 1e-3),
                           "parity": {"n_points": sl, "leaves_identical": bool(np.array_equal(idx[:sl].cpu().numpy(), want))}}
    del v, idx
    # gmm_compute_p: 200k x 64 points, 256 components
    n, d, k = 200_000, 64, 256
    mu = r.rand(k, d).astype(np.float32)
    sg = (0.05 + 0.2 * r.rand(k, d)).astype(np.float32)
    w = r.rand(k).astype(np.float32)
    w /= w.sum()
    g = yl.GmmT(d, k, w.ctypes.data_as(f), mu.ctypes.data_as(f), sg.ctypes.data_as(f))
    v = torch.rand((n, d), device=dev)
    p = torch.empty((n, k), device=dev)
    vp, pp = C.cast(v.data_ptr(), f), C.cast(p.data_ptr(), f)
    L.gmm_compute_p(n, vp, C.byref(g), pp, 1)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        L.gmm_compute_p(n, vp, C.byref(g), pp, 1)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / 3 * 1e3
    sl = 2000
    want = ob.orc_gmm_compute_p(w, mu, sg, v[:sl].cpu().numpy(), 1)
    got = p[:sl].cpu().numpy()
    out["gmm_compute_p"] = {"config": "200k x 64 points, 256 components, GMM_FLAGS_W (device-resident points)",
                            "ms": ms, "tflops_fp32": 4.0 * n * k * d / (ms * 1e-3) / 1e12,
                            "parity": {"n_points": sl, "max_abs_err": float(np.abs(got - want).max()),
                                       "ok": bool(np.abs(got - want).max() <= 1e-5)}}
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
        sh.search(hq)
        L.yb_prof_enable(1)
        L.yb_prof_ms(0, None, 1)
        for _ in range(3):
            dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            sh.search(hq)
            e1.record()
            e1.synchronize()
            best = min(best, max_over_ranks(e0.elapsed_time(e1)))
        ph = _phase_means(L, ((12, "expand_codes"), (13, "sample_thresholds"), (14, "e4m3_pass"),
                              (15, "order_and_certify"), (7, "popcount_scan"), (16, "exchange_x2"),
                              (17, "merge_query_slice")))
        L.yb_prof_ms(0, None, 1)
        L.yb_prof_enable(0)
        out["hamming_knn_10Mx64bit_10kq_k100_sharded"] = {
            "queries_per_s": nqh / (best * 1e-3), "ms": best, "ranks": world,
            "engine": int(L.yb_last_hamming_engine()), "scaling": "strong (10M codes split over the ranks)",
            "phase_ms_rank0": ph,
            "exchange": "peer memory" if (sh.comm and L.yb_comm_p2p(sh.comm.handle)) else "nccl"}
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
        niter = 5
        ph = {}
        for rep in range(2):
            dist.barrier()
            torch.cuda.synchronize()
            if rep == 1:
                L.yb_prof_enable(1)
                L.yb_prof_ms(0, None, 1)
            t = time.perf_counter()
            try:
                _, q, _, _ = ydist.sharded_kmeans(v, k, niter, c0, n)
            except Exception as e:  # the all-reduce hook returns an error code instead of raising
                okrun = False
                out["kmeans_error"] = str(e)
            torch.cuda.synchronize()
            times.append(max_over_ranks((time.perf_counter() - t) / niter))
            if rep == 1:
                ph = _phase_means(L, ((0, "center_and_convert"), (1, "tensor_pass"), (3, "rerank_k1"),
                                      (4, "exact_fallback"), (8, "update_sort_and_sums"), (16, "all_reduce"),
                                      (9, "scale")))
                L.yb_prof_ms(0, None, 1)
                L.yb_prof_enable(0)
            if not all_ok(okrun):
                break
        if okrun and times:
            out["kmeans_10Mx128_k65536_sharded"] = {
                "iter_per_s": 1.0 / min(times), "s_per_iter": min(times), "ranks": world, "qerr": float(q),
                "iterations_timed": niter, "phase_ms_rank0": ph,
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
    poll_ms = float(os.environ.get("BENCH_CLOCK_POLL_MS", "2" if world == 1 else "20"))
    sampler = ClockSampler(local, uuid, poll_ms=poll_ms,
                           enabled=(rank == 0 and os.environ.get("BENCH_CLOCKS", "on") != "off"))
    sampler.start()
    L.yb_launch_count(1)
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
    shard_phase_ms = None
    ms = e0.elapsed_time(e1) / args.steps
    cnt = C.c_long(0)
    phase_ms = {}
    for ph, name in ((0, "center_and_norms"), (10, "sample_thresholds"), (1, "tf32_shortlist"),
                     (2, "merge_select"), (3, "rerank"), (4, "exact_fallback"), (5, "exact_slab"),
                     (6, "row_select"), (16, "exchange_all_to_all_plus_all_gather"), (17, "merge_query_slice")):
        t = L.yb_prof_ms(ph, C.byref(cnt), 0)
        if cnt.value:
            phase_ms[name] = t / cnt.value
    L.yb_prof_ms(0, None, 1)
    L.yb_prof_enable(0)
    if world > 1:
        # the exchange phase is recorded as two spans per search (all-to-all, all-gather)
        if "exchange_all_to_all_plus_all_gather" in phase_ms:
            phase_ms["exchange_all_to_all_plus_all_gather"] *= 2
        shard_phase_ms = {k_: phase_ms[k_] for k_ in ("exchange_all_to_all_plus_all_gather", "merge_query_slice")
                          if k_ in phase_ms}
        # A/B in the same run: the round-1 exchange (all-gather of every list to every rank + full merge)
        ab = ydist.ShardedKnn(base, K, rank=rank, world=world, exchange="torch")
        for _ in range(3):
            ab.search(query)
        dist.barrier()
        torch.cuda.synchronize()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record(stream)
        for _ in range(args.steps):
            ab.search(query)
        a1.record(stream)
        torch.cuda.synchronize()
        tab = torch.tensor([a0.elapsed_time(a1) / args.steps], device=dev)
        dist.all_reduce(tab, op=dist.ReduceOp.MAX)
        shard_phase_ms["ms_per_step_with_round1_allgather_exchange"] = float(tab.item())
        tmax = torch.tensor([ms], device=dev)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        ms = float(tmax.item())
    value = world * NQ / (ms * 1e-3)

    # ---- end to end through the drop-in C call with pinned host buffers
    # N > 1: every rank feeds its own 517 MB shard over its own PCIe link at the same time; allocate the
    # pinned buffers from the NUMA node the GPU hangs off (first touch under a node-local CPU affinity)
    numa = _bind_to_gpu_numa_node(torch, dev) if world > 1 else None
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
    if world > 1:
        e2e["pinned_buffers_numa_node_rank0"] = numa

    # ---- results for the parity checks (outside every timed region)
    L.yb_set_knn_engine(-1)
    loc_i = torch.empty((NQ, K), dtype=torch.int32, device=dev)
    loc_d = torch.empty((NQ, K), dtype=torch.float32, device=dev)
    searcher._local_into(query, loc_i, loc_d, 0)          # this rank's shard alone, local ids
    torch.cuda.synchronize()
    res_idx, res_dis = loc_i.cpu().numpy(), loc_d.cpu().numpy()
    parity_sharded = None
    if world > 1:
        # the merged answer of the sharded search == ONE search of the concatenated shards, for the
        # first 256 queries (every rank gathers all shards over NCCL and checks; rank 0 reports)
        nchk = 256
        mi, md = searcher.search(query)
        allb = torch.empty((world, NB, D), dtype=torch.float32, device=dev)
        dist.all_gather_into_tensor(allb.view(-1), base.view(-1))
        one_i = torch.empty((nchk, K), dtype=torch.int32, device=dev)
        one_d = torch.empty((nchk, K), dtype=torch.float32, device=dev)
        rc = L.yb_knn_l2(nchk, world * NB, D, K, allb.data_ptr(), query.data_ptr(), None, one_i.data_ptr(),
                         one_d.data_ptr(), 0, ydist._stream_ptr(torch))
        torch.cuda.synchronize()
        ok = rc == 0 and bool(torch.equal(one_i, mi[:nchk]) and torch.equal(one_d, md[:nchk]))
        flag = torch.tensor([1 if ok else 0], device=dev, dtype=torch.int32)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        parity_sharded = {"config": "%d shards of 1M x 128 merged vs one search of the %dM-row database" % (world, world),
                          "n_queries": nchk, "ids_and_distances_bit_identical_on_every_rank": bool(flag.item())}
        del allb
        torch.cuda.empty_cache()

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
        traffic = traffic_src = None
        for name in ("r2_knn_tf32_traffic.json", "r1_knn_tf32_traffic.json"):
            try:  # DRAM bytes of this kernel from the committed ncu --set full capture (not re-measured here)
                tj = json.load(open(os.path.join(ROOT, "profiles", name)))
                traffic = tj["dram_bytes_per_launch"]
                traffic_src = "profiles/%s (ncu --set full capture at commit %s)" % (name, tj.get("commit", "see file"))
                break
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
                "unit": "TFLOP/s", "frac": ach / peak, "traffic": traffic, "traffic_source": traffic_src,
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

    cpu, (n_ref, widx, wdis) = cpu_reference_rate(base_h, query_h)
    # parity at the BASELINE size: the reference's answer for the sampled queries against the GPU's
    # (rank 0's shard; for N > 1 the local search of that shard, before the merge)
    parity = knn_parity(res_idx, res_dis, widx, wdis, base_h, query_h,
                        "C2: 1M x 128 database, k=100 (BASELINE configs[1])")

    kmeans_block = hamming_block = cross_block = consumers_block = None
    if world == 1 and not args.no_extras:
        try:
            hamming_block = measure_hamming(torch, L, dev)
        except Exception as e:  # the side blocks never break the headline line
            hamming_block = {"error": str(e)}
        try:
            kmeans_block = measure_kmeans(torch, L, dev, args.kmeans_iters)
        except Exception as e:
            kmeans_block = {"error": str(e)}
        try:
            cross_block = measure_cross(torch, L, dev)
        except Exception as e:
            cross_block = {"error": str(e)}
        try:
            consumers_block = measure_consumers(torch, L, dev)
        except Exception as e:
            consumers_block = {"error": str(e)}
    elif extras_sharded:
        kmeans_block = extras_sharded.get("kmeans_10Mx128_k65536_sharded") or {"error": extras_sharded.get("kmeans_error")}
        hamming_block = extras_sharded.get("hamming_knn_10Mx64bit_10kq_k100_sharded") or {"error": extras_sharded.get("hamming_error")}

    print(json.dumps({
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(3, args.warmup), "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None,
        "dtype": "f32 (%s tensor-core shortlist with FP32 accumulation, exact f32 re-rank)" % operands,
        "data": "synthetic uniform[0,1), seeds 1234/1235",
        "config": {
            "workload": WORKLOAD,
            "parallelism": ("single GPU" if world == 1 else
                            "database sharded x%d (1M rows per rank); query-partitioned exchange inside the "
                            "library (list slices to their owners, merge of nq/N queries per rank, merged slices "
                            "to everybody) over %s; value counts query x 1M-shard scans"
                            % (world, "peer memory (NVLink stores + flag barrier)"
                               if (searcher.comm and L.yb_comm_p2p(searcher.comm.handle)) else "NCCL send/recv")),
            "l2": "database (512 MB) is larger than L2 (126 MB): no flush needed between steps",
            "engine": ("tcgen05 %s + FP32 re-rank" % operands) if engine == 1 else "exact FP32 SIMT",
            "uncertified_queries_redone_exactly": int(uncert),
            "phase_ms": phase_ms,
            "shard_phase_ms_rank0": shard_phase_ms,
        },
        "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
        "roofline": roof, "cpu_baseline": cpu, "parity": parity, "parity_sharded": parity_sharded,
        "kmeans": kmeans_block, "hamming": hamming_block, "cross_distances": cross_block,
        "consumers": consumers_block,
    }))
    if world > 1:
        dist.destroy_process_group()


def run_c5(args):
    """--workload c5: BASELINE configs[4], exact kNN "deep shape" -- 100M x 96 float32 database sharded
    by rows over the ranks (generated on the device, shard by shard), 10 000 queries, k = 100, the
    library's query-partitioned NCCL exchange (reference call shape: progs/knn.c:225).  Prints one
    JSON line: queries/s against the WHOLE database (strong scaling), the tensor pass's roofline
    fraction per rank, and two parity objects on a 1M-row sub-database (1M / N rows of every shard):
    sharded search == ONE search of the concatenation (bit-identical), and the oracle's answer."""
    import torch
    import torch.distributed as dist
    import yael_b200
    from yael_b200 import dist as ydist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    L = yael_b200.lib()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    L.yb_set_device(local)
    if world > 1:
        import datetime
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=600))
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    n_total, d, nq, k = args.c5_rows, 96, 10_000, 100
    lo, hi = ydist.shard_bounds(n_total, world)[rank]
    base = torch.empty((hi - lo, d), dtype=torch.float32, device=dev)
    g = torch.Generator(device=dev)
    step = 4_000_000
    for a in range(0, hi - lo, step):   # seeded per 4M-row block of the GLOBAL database: same data for any N
        b = min(hi - lo, a + step)
        base[a:b].uniform_(0, 1, generator=g.manual_seed(4321 + (lo + a) // 1000))
    gq = torch.Generator(device=dev)
    gq.manual_seed(77)
    query = torch.rand((nq, d), device=dev, dtype=torch.float32, generator=gq)
    searcher = ydist.ShardedKnn(base, k, rank=rank, world=world, id_offset=lo)
    for _ in range(max(2, args.warmup)):
        searcher.search(query)
    torch.cuda.synchronize()
    L.yb_prof_enable(1)
    L.yb_prof_ms(1, None, 1)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    steps = max(2, min(args.steps, 5))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        res_i, res_d = searcher.search(query)
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    cnt = C.c_long(0)
    phase_ms = {}
    for ph, name in ((0, "center_and_norms"), (10, "sample_thresholds"), (1, "tf32_shortlist"), (2, "merge_select"),
                     (3, "rerank"), (4, "exact_fallback"), (16, "exchange"), (17, "merge_query_slice")):
        t = L.yb_prof_ms(ph, C.byref(cnt), 0)
        if cnt.value:
            phase_ms[name] = t / cnt.value * (2 if ph == 16 else 1)
    L.yb_prof_ms(0, None, 1)
    L.yb_prof_enable(0)
    uncert = int(L.yb_last_knn_uncertified())
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    # ---- parity on a 1M-row sub-database: sub_n / N rows of every shard
    sub = args.c5_sub // world
    sub_base = base[:sub].contiguous()
    nchk = 512
    s_sub = ydist.ShardedKnn(sub_base, k, rank=rank, world=world, id_offset=rank * sub)
    mi, md = s_sub.search(query[:nchk].contiguous())
    if world > 1:
        allb = torch.empty((world * sub, d), dtype=torch.float32, device=dev)
        dist.all_gather_into_tensor(allb.view(-1), sub_base.view(-1))
    else:
        allb = sub_base
    one_i = torch.empty((nchk, k), dtype=torch.int32, device=dev)
    one_d = torch.empty((nchk, k), dtype=torch.float32, device=dev)
    rc = L.yb_knn_l2(nchk, world * sub, d, k, allb.data_ptr(), query.data_ptr(), None, one_i.data_ptr(),
                     one_d.data_ptr(), 0, ydist._stream_ptr(torch))
    torch.cuda.synchronize()
    same = rc == 0 and bool(torch.equal(one_i, mi) and torch.equal(one_d, md))
    if world > 1:
        f = torch.tensor([1 if same else 0], device=dev, dtype=torch.int32)
        dist.all_reduce(f, op=dist.ReduceOp.MIN)
        same = bool(f.item())
    if rank == 0:
        from oracle import bindings as ob
        nor = 64
        bh, qh = allb.cpu().numpy(), query[:nor].cpu().numpy()
        cores = len(os.sched_getaffinity(0))
        widx, wdis = (ob.ref_knn(bh, qh, k, nt=cores) if ob.have_ref() else
                      ob.orc_knn(bh, qh, k, dot_mode=ob.DOT_F32_SEQ, nt=cores))
        par = knn_parity(mi[:nor].cpu().numpy(), md[:nor].cpu().numpy(), widx, wdis, bh, qh,
                         "C5 sub-database: %d x 96 (%d rows of each of the %d shards), k=100" % (world * sub, sub, world))
        peaks = _peaks()
        bf16 = peaks.get("bf16_tflops", 1590.0)
        roof = None
        if "tf32_shortlist" in phase_ms:
            ach = 2.0 * nq * (hi - lo) * d / (phase_ms["tf32_shortlist"] * 1e-3) / 1e12
            roof = {"bound": "tensor", "kernel": "k_knn_tf32<EPI_LISTS, fp16> on this rank's shard",
                    "achieved": ach, "peak": bf16, "unit": "TFLOP/s", "frac": ach / bf16, "traffic": None,
                    "kernel_ms": phase_ms["tf32_shortlist"]}
        print(json.dumps({
            "metric": "kNN queries/s (%dM x 96 db sharded over %d GPUs, k=100)" % (n_total // 1_000_000, world),
            "value": nq / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": steps, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "strong", "data": "synthetic uniform[0,1), generated on the device",
            "config": {"workload": "exact kNN, deep shape: %d x 96 database sharded by rows over %d GPUs, 10000 "
                                   "queries, k=100, NCCL top-k merge (BASELINE configs[4])" % (n_total, world),
                       "rows_per_rank": hi - lo, "phase_ms_rank0": phase_ms,
                       "uncertified_queries_redone_exactly": uncert,
                       "operands": "fp16" if L.yb_last_knn_operands() == 2 else "tf32"},
            "roofline": roof,
            "parity_sharded_vs_one_search": {"n_queries": nchk, "rows": world * sub,
                                             "ids_and_distances_bit_identical_on_every_rank": same},
            "parity": par}))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-extras", action="store_true",
                    help="skip the k-means / Hamming blocks")
    ap.add_argument("--kmeans-iters", type=int, default=10,
                    help="iterations of the k-means block (BASELINE configs[3] names 10)")
    ap.add_argument("--workload", default="c2", choices=["c2", "c5"],
                    help="c2: the headline (BASELINE configs[1]); c5: 100M x 96 sharded (configs[4])")
    ap.add_argument("--c5-rows", type=int, default=100_000_000)
    ap.add_argument("--c5-sub", type=int, default=1_000_000)
    args = ap.parse_args()
    if args.workload == "c5" and args.impl == "ours":
        run_c5(args)
        return
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
