#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native EmbeddingNet hot path (contract: see DESIGN.md section 6).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--skip-knn] [--knn-bank ROWS]

Primary metric (BASELINE.json): fused batch-hard triplet loss + gradient, embeddings/sec at B = 4096, d = 512
(config C3; the in-batch path does not shard: N > 1 runs N independent replicas, "weak" scaling).
Secondary, on the same JSON line under "knn": kNN queries/sec, 100k queries vs a 10M x 512 fp32 bank, k = 5, the
bank sharded row-wise over the N ranks and merged after one NCCL all-gather -- the only partitioned path.

One "step" = one pass of the hot path over one batch of synthetic input (counter-hash generator, bit-identical on
CPU and GPU).  `value` is device-resident throughput (CUDA events on the launching stream, L2 flushed between
steps, max over ranks); `e2e` goes through the reference-shaped public API with pinned HOST buffers and includes
the host<->device copies.  `--impl reference` times the reference's CPU path (oracle port: scikit-learn + NumPy,
all host threads) for the same metric and config; TensorFlow 2.2 itself cannot run in this image.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

B, D, N_CLASSES, PER_CLASS, MARGIN = 4096, 512, 512, 8, 0.5
KNN_Q, KNN_BANK, KNN_K, KNN_CLASSES = 100_000, 10_000_000, 5, 100_000
TRIPLET_WORKLOAD = ("C3: fused batch-hard triplet loss + grad, B=4096 d=512 (512 classes x 8, post-ReLU, "
                    "L2-normalised), margin 0.5, non-squared distances")


def bench_config(world):
    """Identical in both arms (the driver compares the two `config` dicts)."""
    return {"workload": TRIPLET_WORKLOAD,
            "parallelism": "replicas only (the in-batch path does not shard)" if world > 1 else "1 GPU"}


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(hbm_gbs=float(p["hbm_gbs"]), bf16=float(p["bf16_tflops"]),
                    bf16_sustained=float(p.get("bf16_tflops_sustained", p["bf16_tflops"])), source="measured")
    return dict(hbm_gbs=6650.0, bf16=1590.0, bf16_sustained=1400.0, source="fallback")


def ncu_traffic(key):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture (profiles/), or None."""
    path = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(path):
        try:
            with open(path) as f:
                return json.load(f).get(key)
        except Exception:
            return None
    return None


# ----------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples SM clock and throttle reasons through NVML while the timed regions run."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4),
        }
        while not self._stop.is_set():
            try:
                util = nv.nvmlDeviceGetUtilizationRates(self.h).gpu
                mhz = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                if util > 0:
                    self.samples.append(mhz)
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.02)

    def start(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()

    def stop(self):
        self._stop.set()
        if self._thr is not None:
            self._thr.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ----------------------------------------------------------------------------------------------- reference arm
def triplet_inputs_numpy():
    from embeddingnet_b200 import synth

    x, lab = synth.make_numpy(B, D, n_classes=N_CLASSES, rows_per_class=PER_CLASS, noise=0.5, relu=True)
    import numpy as np

    ss = np.sum(x.astype(np.float64) ** 2, axis=1, keepdims=True)
    x = (x / np.sqrt(np.maximum(ss, 1e-12))).astype(np.float32)
    return x, lab


def cpu_triplet_baseline(steps, warmup):
    """Reference arm of the headline metric: sklearn.pairwise_distances (the reference's own distance call) +
    NumPy batch-hard selection and gradient, on all host threads."""
    from oracle import np_oracle as O

    x, lab = triplet_inputs_numpy()
    for _ in range(warmup):
        O.batch_hard_loss_grad_cpu(lab, x, MARGIN)
    t0 = time.perf_counter()
    for _ in range(steps):
        O.batch_hard_loss_grad_cpu(lab, x, MARGIN)
    dt = (time.perf_counter() - t0) / steps
    return B / dt, dt


def cpu_knn_baseline(bank_rows=400_000, n_q=256):
    """Brute-force KNeighborsClassifier (what models.py:136-138 expects) on a bounded sample; brute kNN is linear
    in bank rows, so q/s against the full bank = measured q/s * bank_rows / 10M."""
    from sklearn.neighbors import KNeighborsClassifier

    from embeddingnet_b200 import synth

    bank, labels = synth.make_numpy(bank_rows, D, n_classes=KNN_CLASSES, noise=0.5)
    q, _ = synth.make_numpy(n_q, D, seed_noise=synth.SEED_QUERY, n_classes=KNN_CLASSES, noise=0.5)
    clf = KNeighborsClassifier(n_neighbors=KNN_K, algorithm="brute").fit(bank, labels)
    clf.kneighbors(q[:8])
    t0 = time.perf_counter()
    clf.kneighbors(q, n_neighbors=KNN_K)
    dt = time.perf_counter() - t0
    qps_sample = n_q / dt
    return qps_sample * bank_rows / KNN_BANK, "%d queries vs a %d-row slice of the bank, scaled linearly to 10M rows" % (
        n_q, bank_rows)


def use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1 to every rank; the reference arm runs on rank 0 alone and is meant to use
    all host cores.  Must run before NumPy / scikit-learn load their BLAS."""
    n = str(os.cpu_count() or 1)
    for var in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS", "NUMEXPR_NUM_THREADS"):
        os.environ[var] = n


def run_reference(args, rank):
    if rank != 0:
        return
    cores = os.cpu_count()
    try:
        from threadpoolctl import threadpool_info, threadpool_limits

        threadpool_limits(limits=cores)
        import numpy  # noqa: F401  (loads the BLAS so that threadpool_info sees it)

        threads = max([int(p.get("num_threads", 1)) for p in threadpool_info()] or [1])
    except Exception:
        threads = cores
    # same K and W as our arm (one C3 step takes ~0.2 s on the host: 50 steps stay within seconds)
    W, steps = max(3, args.warmup), max(1, args.steps)
    value, dt = cpu_triplet_baseline(steps, W)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    line = {
        "impl": "reference",
        "metric": "batch_hard_triplet_loss_grad_embeddings_per_sec", "value": value, "unit": "embeddings/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": W, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32 in / f64 BLAS internals",
        "data": "synthetic",
        "config": bench_config(world),
        "method": "reference CPU path: sklearn.pairwise_distances (datagenerators.py:219) + NumPy selection and "
                  "gradient on rank 0's host cores; TensorFlow 2.2 is not runnable in this image (Python 3.12, no "
                  "network)",
        "cpu_baseline": {"value": value, "unit": "embeddings/s", "cores": threads, "kind": "port",
                         "sample": "full C3 step (B=4096, d=512), %d steps, %d BLAS threads of %d host cores" % (
                             steps, threads, cores)},
        "e2e": {"value": value, "unit": "embeddings/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if not args.skip_knn:
        qps, sample = cpu_knn_baseline()
        line["knn"] = {"metric": "knn_queries_per_sec_10M_bank", "value": qps, "unit": "queries/s",
                       "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": cores, "kind": "port",
                                        "sample": sample}}
    print(json.dumps(line), flush=True)


def rowwise_rooflines(torch, _lib, synth, dev, flush, peaks):
    """HBM roofline of the memory-bound row-wise kernels (north_star (3); SURVEY 8(d) A6 / A9 / A10 byte counts),
    called through the C ABI.  C3 scale (4096 x 512: 8-25 MB, smaller than L2, so L2 is flushed before every timed
    launch and the figure includes launch latency) and bank scale (1M x 256 = 1 GB per operand, >> L2)."""
    from embeddingnet_b200._runtime import ptr, stream_ptr

    out = {}
    for tag, rows, d, flush_each in (("C3_4096x512", 4096, 512, True), ("bank_1Mx256", 1_000_000, 256, False)):
        x = synth.make_device(rows, d, n_classes=max(rows // 8, 1), rows_per_class=8, noise=0.5, relu=True, device=dev)[0]
        g = synth.make_device(rows, d, seed_noise=77, device=dev)[0]
        y = torch.empty_like(x)
        apn = synth.make_device(rows, 3 * d, seed_noise=78, device=dev)[0]
        gapn = torch.empty_like(apn)
        vec = torch.empty(rows, dtype=torch.float32, device=dev)
        ones = torch.ones(rows, dtype=torch.float32, device=dev)
        s = stream_ptr()
        m = ctypes.c_float(0.5)
        n4 = rows * d * 4
        cases = {
            "l2_normalize_fwd": (2 * n4, lambda: _lib.call("en_l2_normalize_fwd", ptr(x), ptr(y), rows, d, s)),
            "l2_normalize_bwd": (3 * n4, lambda: _lib.call("en_l2_normalize_bwd", ptr(x), ptr(g), ptr(y), rows, d, s)),
            "triplet_apn_fwd": (3 * n4 + rows * 4, lambda: _lib.call("en_triplet_apn_fwd", ptr(apn), rows, 3 * d, m,
                                                                     ptr(vec), s)),
            "triplet_apn_bwd": (6 * n4 + rows * 4, lambda: _lib.call("en_triplet_apn_bwd", ptr(apn), ptr(ones), rows,
                                                                     3 * d, m, ptr(gapn), s)),
            "siamese_l2_fwd": (2 * n4 + rows * 4, lambda: _lib.call("en_siamese_l2_fwd", ptr(x), ptr(g), rows, d,
                                                                    ptr(vec), s)),
            "query_distances": (n4 + rows * 4, lambda: _lib.call("en_query_distances", ptr(x), ptr(g), rows, d,
                                                                 ptr(vec), s)),
        }
        rec = {}
        for name, (nbytes, fn) in cases.items():
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            if flush_each:
                evs = []
                for _ in range(20):
                    flush.zero_()
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record()
                    fn()
                    b.record()
                    evs.append((a, b))
                torch.cuda.synchronize()
                ms = sorted(a.elapsed_time(b) for a, b in evs)[len(evs) // 2]
            else:
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                for _ in range(10):
                    fn()
                b.record()
                torch.cuda.synchronize()
                ms = a.elapsed_time(b) / 10
            gbs = nbytes / (ms * 1e-3) / 1e9
            rec[name] = {"ms": ms, "algorithmic_bytes": nbytes, "achieved_gbs": gbs, "frac": gbs / peaks["hbm_gbs"]}
        out[tag] = rec
        del x, g, y, apn, gapn, vec, ones
        torch.cuda.empty_cache()
    out["peak_gbs"] = peaks["hbm_gbs"]
    out["note"] = ("C3-scale launches move 8-50 MB in 3-10 us: launch latency and the L2-flushed cold start are inside "
                   "the figure; the bank-scale rows are the steady-state HBM rate")
    return out


# ----------------------------------------------------------------------------------------------- ours
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--skip-knn", action="store_true")
    ap.add_argument("--knn-bank", type=int, default=KNN_BANK)
    ap.add_argument("--knn-queries", type=int, default=KNN_Q)
    ap.add_argument("--knn-steps", type=int, default=2)
    ap.add_argument("--skip-cpu", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        if rank == 0:
            use_all_host_threads()
        run_reference(args, rank)
        return

    import numpy as np
    import torch

    from embeddingnet_b200 import _lib, synth
    from embeddingnet_b200 import losses_and_accuracies as lac
    from embeddingnet_b200._runtime import launch_count, launch_count_reset
    from embeddingnet_b200.fused import BatchHardStep
    from embeddingnet_b200.models import BankKNNClassifier

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device; there is no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    group = None
    if world > 1:
        import torch.distributed as dist

        # NCCL's own init log (rank / nranks lines) is the evidence that the sharded path really spans N ranks: keep
        # it on, but on stderr, so that stdout stays the one JSON line
        # (through a per-rank file that is replayed on stderr at the end: pointing NCCL_DEBUG_FILE at /dev/stderr makes
        # every rank fopen(.., "w") -- i.e. truncate -- whatever file stderr is redirected to)
        # (the image presets NCCL_DEBUG=VERSION, which prints a banner on stdout: raised to INFO like an unset value; a
        # caller's INFO / TRACE is left as it is)
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() in ("VERSION", "WARN"):
            os.environ["NCCL_DEBUG"] = "INFO"
            os.environ.setdefault("NCCL_DEBUG_SUBSYS", "INIT")
        nccl_log = None
        if "NCCL_DEBUG_FILE" not in os.environ:
            import tempfile

            nccl_log = os.path.join(tempfile.gettempdir(), "en_bench_nccl_%s_r%d.log" % (
                os.environ.get("MASTER_PORT", "0"), rank))
            os.environ["NCCL_DEBUG_FILE"] = nccl_log
        dist.init_process_group("nccl", device_id=dev)
        group = dist.group.WORLD
    lib = _lib.load()  # raises if the CUDA extension is missing
    peaks = load_peaks()
    W, K = max(3, args.warmup), max(1, args.steps)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    sampler = ClockSampler(local_rank)
    sampler.start()

    # ------------------------------------------------------------------ primary: batch-hard loss + grad
    raw, labels = synth.make_device(B, D, n_classes=N_CLASSES, rows_per_class=PER_CLASS, noise=0.5, relu=True, device=dev)
    emb = lac.l2_normalize(raw).detach().contiguous()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2
    stepper = BatchHardStep(B, D, margin=MARGIN)
    stepper.step(emb, labels)
    torch.cuda.synchronize()
    # the whole loss+grad step as one CUDA graph: launch-bound inner loop, no host work between kernels
    graph = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        stepper.step(emb, labels)
        side.synchronize()
        launch_count_reset()
        with torch.cuda.graph(graph, stream=side):
            stepper.step(emb, labels)
        launches_per_step = launch_count()
    torch.cuda.current_stream().wait_stream(side)
    for _ in range(W):
        flush.zero_()
        graph.replay()
    barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    for i in range(K):
        flush.zero_()
        ev[i][0].record()
        graph.replay()
        ev[i][1].record()
    barrier()
    step_ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = max_over_ranks(sum(step_ms))
    ms_per_step = total_ms / K
    value = world * B * K / (total_ms * 1e-3)
    loss_value = float(stepper.loss.item())

    # dominant kernel (distance GEMM) timed live with CUDA events on its launching stream
    lib.en_prof_enable(1)
    gemm_ms = []
    for _ in range(10):
        flush.zero_()
        stepper.step(emb, labels)
        ms = ctypes.c_float(0)
        _lib.check(lib.en_prof_last_ms(ctypes.byref(ms)), "en_prof_last_ms")
        gemm_ms.append(ms.value)
    lib.en_prof_enable(0)
    gemm_ms_avg = sum(gemm_ms[2:]) / len(gemm_ms[2:])
    flops = 2.0 * B * B * D
    achieved = flops / (gemm_ms_avg * 1e-3) / 1e12
    # the batch-hard GEMM runs on split-BF16 operands (3 kind::f16 MMAs per k-step): its own tensor roofline is
    # dense bf16 / 3; north_star's "TF32-emulated roofline" (dense bf16 / 2 / 3) is reported beside it
    peak = peaks["bf16"] / 3.0
    n_t = B // 128
    tiles_exec = 2 * sum((n_t + 1) // 2 - (i >> 1) for i in range(n_t))   # csrc/tc_engine_wide.cuh: num_items() pairs
    roofline = {
        "bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
        "traffic": ncu_traffic("batch_hard_gemm_dram_bytes_per_launch"),
        "kernel": "dist_gemm_wide_kernel<EpBatchHard> (tcgen05 kind::f16 on split-BF16 planes, 3 N=256 MMAs per "
                  "k-step; work item = row tile x pair of column tiles)",
        "kernel_ms": gemm_ms_avg, "share_of_step": gemm_ms_avg / ms_per_step,
        "algorithmic_flops_per_launch": flops,
        "peak_note": "%s bf16 dense %.1f TFLOP/s (burst) / 3 (hi*hi + hi*lo + lo*hi passes)" % (
            peaks["source"], peaks["bf16"]),
        "frac_of_bf16_peak": achieved / peaks["bf16"],
        "frac_of_tf32x3_roofline": achieved / (peaks["bf16"] / 6.0),
        # `achieved` credits the ALGORITHMIC 2 B^2 d flops.  The symmetric schedule executes only the
        # tile pairs that touch the upper triangle (272 pairs = 544 tiles of 1024 at B = 4096), three kind::f16 MMAs
        # per k-step each: the rate the tensor pipe actually runs at is below, so `frac` (which can reach
        # 1024/544 = 1.88 for this schedule) is not misread as pipe utilisation.  What bounds the kernel is the
        # L2 -> SM operand stream (tools/trace_bh.py), reported as l2_operand_stream_tbs.
        "tiles_executed": tiles_exec, "tiles_algorithmic": (B // 128) ** 2,
        "hardware_mma_tflops": 3.0 * tiles_exec * 2.0 * 128 * 128 * D / (gemm_ms_avg * 1e-3) / 1e12,
        "hardware_frac_of_bf16_peak": 3.0 * tiles_exec * 2.0 * 128 * 128 * D / (gemm_ms_avg * 1e-3) / 1e12 / peaks["bf16"],
        "l2_operand_stream_tbs": (tiles_exec // 2) * (D // 64) * 96 * 1024 / (gemm_ms_avg * 1e-3) / 1e12,
    }

    # end to end through the reference-shaped public API, host buffers in and out
    emb_h = emb.cpu().pin_memory()
    lab_h = labels.cpu().pin_memory()
    grad_h = torch.empty((B, D), dtype=torch.float32).pin_memory()
    fn = lac.batch_hard_triplet_loss(MARGIN)

    def e2e_step():
        e = emb_h.to(dev, non_blocking=True).requires_grad_(True)
        l = lab_h.to(dev, non_blocking=True)
        loss = fn(l, e)
        loss.backward()
        grad_h.copy_(e.grad, non_blocking=True)
        return float(loss.item())  # device -> host read of the step's result (synchronises)

    for _ in range(W):
        e2e_step()
    barrier()
    launch_count_reset()
    t0 = time.perf_counter()
    for _ in range(K):
        e2e_loss = e2e_step()
    torch.cuda.synchronize()
    e2e_serial_dt = max_over_ranks(time.perf_counter() - t0)
    e2e_launches = launch_count() // K
    assert abs(e2e_loss - loss_value) <= 1e-6 * max(1.0, abs(loss_value)), (e2e_loss, loss_value)

    # The same work as a 3-deep software pipeline (what a data-loader-fed training loop does): step i's host->device
    # copy, step i-1's compute and step i-2's device->host copies run on three streams; every step still moves its
    # own inputs in and its own gradient + loss out, and every loss is read on the host inside the timed region.
    s_in, s_cmp, s_out = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()
    DEPTH = 3
    e_dev = [torch.empty((B, D), dtype=torch.float32, device=dev) for _ in range(DEPTH)]
    l_dev = [torch.empty(B, dtype=torch.int32, device=dev) for _ in range(DEPTH)]
    g_host = [torch.empty((B, D), dtype=torch.float32).pin_memory() for _ in range(DEPTH)]
    l_host = [torch.empty((), dtype=torch.float32).pin_memory() for _ in range(DEPTH)]
    ev_in = [torch.cuda.Event() for _ in range(DEPTH)]
    ev_cmp = [torch.cuda.Event() for _ in range(DEPTH)]
    ev_out = [torch.cuda.Event() for _ in range(DEPTH)]
    keep = [None] * DEPTH
    steppers = [BatchHardStep(B, D, margin=MARGIN) for _ in range(DEPTH)]  # one per slot: each owns its gradient

    def compute_autograd(k):
        e = e_dev[k].detach().requires_grad_(True)
        loss = fn(l_dev[k], e)
        loss.backward()
        return loss.detach(), e.grad

    def compute_cabi(k):
        return steppers[k].step(e_dev[k], l_dev[k])  # en_batch_hard_fwd_bwd: loss and gradient from one C-ABI call

    def pipelined(n_steps, compute):
        losses = []
        for i in range(n_steps + DEPTH - 1):
            if i < n_steps:
                k = i % DEPTH
                if i >= DEPTH:   # slot reuse: the step that used it has been read back (see below)
                    s_in.wait_event(ev_out[k])
                with torch.cuda.stream(s_in):
                    e_dev[k].copy_(emb_h, non_blocking=True)
                    l_dev[k].copy_(lab_h, non_blocking=True)
                    ev_in[k].record()
                with torch.cuda.stream(s_cmp):
                    s_cmp.wait_event(ev_in[k])
                    if i >= DEPTH:
                        s_cmp.wait_event(ev_out[k])  # the slot's previous gradient has left the device
                    loss, grad = compute(k)
                    ev_cmp[k].record()
                with torch.cuda.stream(s_out):
                    s_out.wait_event(ev_cmp[k])
                    g_host[k].copy_(grad, non_blocking=True)
                    l_host[k].copy_(loss, non_blocking=True)
                    ev_out[k].record()
                keep[k] = (loss, grad)  # alive until their copies have been read
            j = i - (DEPTH - 1)
            if j >= 0:           # read the result of step j on the host
                kj = j % DEPTH
                ev_out[kj].synchronize()
                losses.append(float(l_host[kj]))
        return losses

    def timed_pipeline(compute):
        torch.cuda.synchronize()
        pipelined(W, compute)
        barrier()
        t0 = time.perf_counter()
        pl = pipelined(K, compute)
        torch.cuda.synchronize()
        dt = max_over_ranks(time.perf_counter() - t0)
        assert len(pl) == K and all(abs(x - loss_value) <= 1e-6 * max(1.0, abs(loss_value)) for x in pl), pl[:4]
        return dt

    e2e_auto_dt = timed_pipeline(compute_autograd)
    e2e_py_dt = timed_pipeline(compute_cabi)

    # The same work inside the library: en_bh_host_pipe_submit / _wait take HOST pointers (the reference-facing call
    # for a loop that owns host buffers) and make the stage hand-offs on the calling thread (csrc/host_pipe.cu).
    from embeddingnet_b200.fused import BatchHardHostPipeline
    PD = 5  # slots; results are read four steps behind the newest submit
    pipe = BatchHardHostPipeline(B, D, margin=MARGIN, depth=PD)
    p_grad = [BatchHardHostPipeline.pinned((B, D)) for _ in range(PD)]
    p_loss = [BatchHardHostPipeline.pinned((1,)) for _ in range(PD)]

    def host_pipe(n_steps):
        losses = []
        for i in range(n_steps + PD - 1):
            if i < n_steps:
                k = i % PD
                pipe.submit(emb_h, lab_h, p_loss[k], p_grad[k])   # ticket == submit index (checked below)
            j = i - (PD - 1)
            if j >= 0:          # every step's loss is read on the host, in order, inside the timed region
                pipe.wait(host_pipe.base + j)
                losses.append(float(p_loss[j % PD][0]))
        host_pipe.base += n_steps
        return losses

    host_pipe.base = 0
    torch.cuda.synchronize()
    host_pipe(W)
    barrier()
    launch_count_reset()
    t0 = time.perf_counter()
    pl = host_pipe(K)
    torch.cuda.synchronize()
    e2e_dt = max_over_ranks(time.perf_counter() - t0)
    e2e_launches = launch_count() // K
    assert len(pl) == K and all(abs(x - loss_value) <= 1e-6 * max(1.0, abs(loss_value)) for x in pl), pl[:4]
    g_ref = stepper.step(emb, labels)[1]
    assert torch.equal(p_grad[(K - 1) % PD].to(dev), g_ref) or \
        (p_grad[(K - 1) % PD].to(dev) - g_ref).norm() <= 1e-6 * g_ref.norm()  # atomics: summation order only
    pipe.close()
    e2e = {"value": world * B * K / e2e_dt, "unit": "embeddings/s", "h2d_bytes_per_step": B * D * 4 + B * 4,
           "d2h_bytes_per_step": B * D * 4 + 4, "ms_per_step": e2e_dt / K * 1e3,
           "api": "en_bh_host_pipe_submit / en_bh_host_pipe_wait (C ABI, HOST pointers: pinned embeddings + labels "
                  "in, loss + gradient out) through fused.BatchHardHostPipeline",
           "schedule": "5 slots inside the library over three CUDA streams (H2D | one CUDA graph of the 4 step kernels | "
                       "D2H), stage hand-offs made by the calling thread inside submit / wait (no stream waits on "
                       "another stream's event); every step copies its own inputs in and its gradient + loss out, "
                       "every loss is read on the host in order, four steps behind the newest submit",
           "python_pipeline": {"value": world * B * K / e2e_py_dt, "ms_per_step": e2e_py_dt / K * 1e3,
                               "api": "fused.BatchHardStep.step (en_batch_hard_fwd_bwd on device buffers) with the "
                                      "same 3-stream schedule written in Python (host-bound: ~0.19 ms of "
                                      "interpreter + driver calls per step)"},
           "autograd_api": {"value": world * B * K / e2e_auto_dt, "ms_per_step": e2e_auto_dt / K * 1e3,
                            "api": "losses_and_accuracies.batch_hard_triplet_loss(0.5)(labels, emb); loss.backward() "
                                   "(the reference-shaped callable; same pipeline, ~0.2 ms of Python per step)"},
           "serial": {"value": world * B * K / e2e_serial_dt, "ms_per_step": e2e_serial_dt / K * 1e3,
                      "note": "the reference-shaped callable, strictly one call after the other "
                              "(copy in, compute, copy out, read)"}}

    line = {
        "metric": "batch_hard_triplet_loss_grad_embeddings_per_sec", "value": value, "unit": "embeddings/s",
        "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None,
        "dtype": "f32 in/out; tensor-core selection on split-BF16 planes (3 MMAs per k-step), float64 re-evaluation "
                 "of every selected distance",
        "data": "synthetic",
        "config": bench_config(world),
        "method": {"l2": "256 MiB memset between timed steps (L2 flush)",
                   "timing": "per-step CUDA events around one CUDA-graph replay of fwd+bwd; sum over steps, max over ranks",
                   "loss": loss_value},
        "roofline": roofline, "e2e": e2e, "gpu_launches": int(launches_per_step * K),
        "gpu_launches_per_step": int(launches_per_step), "e2e_gpu_launches_per_step": int(e2e_launches),
    }

    # ------------------------------------------------------------------ the other BASELINE configs (parity-test
    # cases, timed here for the record; rank-local, not part of `value`)
    def time_ms(fn, n=10, w=3):
        for _ in range(w):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / n

    def fwd_bwd(loss_fn, x):
        e = x.detach().clone().requires_grad_(True)
        loss_fn(labels, e).backward()
        return e.grad

    others = {}
    if rank == 0:
        ba = lac.batch_all_triplet_loss(MARGIN, max_positives=PER_CLASS - 1)
        ca = lac.contrastive_loss_all_pairs()
        emb7 = (emb * 0.7).contiguous()
        t_ba = time_ms(lambda: fwd_bwd(ba, emb))
        t_ca = time_ms(lambda: fwd_bwd(ca, emb7))
        pair_flops = 4.0 * B * B * D   # forward distance GEMM + backward (C . E) contraction, 2 B^2 d each
        pair_peak = peaks["bf16"] / 6.0  # 3xTF32: dense TF32 = dense bf16 / 2, three passes

        def pair_roofline(ms):
            ach = pair_flops / (ms * 1e-3) / 1e12
            return {"bound": "tensor", "achieved": ach, "peak": pair_peak, "unit": "TFLOP/s", "frac": ach / pair_peak,
                    "algorithmic_flops_per_step": pair_flops,
                    "peak_note": "%s bf16 dense %.1f TFLOP/s / 2 (TF32) / 3 (passes); whole fwd+bwd step, all "
                                 "kernels" % (peaks["source"], peaks["bf16"])}

        others["C3_batch_all_loss_grad"] = {"ms": t_ba, "embeddings_per_sec": B / (t_ba * 1e-3),
                                            "roofline": pair_roofline(t_ba)}
        others["C3_contrastive_all_pairs_loss_grad"] = {"ms": t_ca, "embeddings_per_sec": B / (t_ca * 1e-3),
                                                        "roofline": pair_roofline(t_ca)}
        # stress shape of SURVEY 8(d): 64 classes x 64 rows (63 positives per anchor)
        raw64, lab64 = synth.make_device(B, D, n_classes=64, rows_per_class=64, noise=0.5, relu=True, device=dev)
        emb64 = lac.l2_normalize(raw64).detach().contiguous()
        ba64 = lac.batch_all_triplet_loss(MARGIN, max_positives=63)

        def fwd_bwd64():
            e = emb64.detach().clone().requires_grad_(True)
            ba64(lab64, e).backward()

        t_ba64 = time_ms(fwd_bwd64, n=5, w=2)
        others["C3_batch_all_loss_grad_64x64"] = {"ms": t_ba64, "embeddings_per_sec": B / (t_ba64 * 1e-3),
                                                  "roofline": pair_roofline(t_ba64)}
        del raw64, emb64
        others["rowwise_hbm"] = rowwise_rooflines(torch, _lib, synth, dev, flush, peaks)
        # C1: reference-semantics in-batch mining, 32 classes x 8 samples, d = 128 (host arrays in, triplets out)
        from embeddingnet_b200.datagenerators import mine_batch_triplets

        x1, l1 = synth.make_numpy(256, 128, n_classes=32, rows_per_class=8, noise=0.5, relu=True)
        x1 = x1 / np.sqrt(np.maximum((x1.astype(np.float64) ** 2).sum(1, keepdims=True), 1e-12)).astype(np.float32)
        np.random.seed(0)
        t0 = time.perf_counter()
        for _ in range(20):
            trip, _ = mine_batch_triplets(x1, l1, margin=MARGIN, mode="semihard")
        t_c1 = (time.perf_counter() - t0) / 20
        others["C1_mining_semihard_256x128_host_to_triplets"] = {"ms": t_c1 * 1e3, "triplets": int(len(trip)),
                                                                 "embeddings_per_sec": 256 / t_c1}
        # C2-shaped step: B = 128, d = 256 batch-hard loss + grad (launch-latency bound)
        x2, l2 = synth.make_device(128, 256, n_classes=16, rows_per_class=8, noise=0.5, relu=True, device=dev)
        x2 = lac.l2_normalize(x2).detach()
        st2 = BatchHardStep(128, 256, margin=MARGIN)
        t_c2 = time_ms(lambda: st2.step(x2, l2), n=50)
        others["C2_batch_hard_loss_grad_128x256"] = {"ms": t_c2, "embeddings_per_sec": 128 / (t_c2 * 1e-3)}
    line["other_configs"] = others

    # ------------------------------------------------------------------ CPU baseline (rank 0, N = 1 only)
    if rank == 0 and world == 1 and not args.skip_cpu:
        cpu_value, cpu_dt = cpu_triplet_baseline(30, 2)
        line["cpu_baseline"] = {"value": cpu_value, "unit": "embeddings/s", "cores": os.cpu_count(), "kind": "port",
                                "sample": "full C3 step on the host, 30 steps (sklearn.pairwise_distances + NumPy "
                                          "selection/gradient); TF 2.2 not runnable here"}

    # ------------------------------------------------------------------ secondary: sharded bank kNN
    if not args.skip_knn:
        del flush
        torch.cuda.empty_cache()
        n_total, Q = int(args.knn_bank), int(args.knn_queries)
        lo, hi = BankKNNClassifier.shard_bounds(n_total, world, rank)
        bank, _ = synth.make_device(hi - lo, D, row_offset=lo, n_classes=KNN_CLASSES, noise=0.5, device=dev)
        label_ids = (torch.arange(n_total, dtype=torch.int64, device=dev) % KNN_CLASSES).to(torch.int32)
        clf = BankKNNClassifier(n_neighbors=KNN_K, process_group=group, device=dev)
        clf.fit_shard(bank, label_ids, lo, n_total, classes=np.arange(KNN_CLASSES))
        # 90 % of the queries are bank rows perturbed by 0.25 u (the true nearest row is known: the
        # planted one), 10 % are unrelated draws.  Bank rows come from the counter-hash generator, so every rank can
        # regenerate the planted rows wherever they live: queries are bit-identical for every N.
        queries, _ = synth.make_device(Q, D, seed_noise=synth.SEED_QUERY, n_classes=KNN_CLASSES, noise=0.5, device=dev)
        n_planted = Q - Q // 10
        blk = 1000
        n_blk = (n_planted + blk - 1) // blk
        planted = torch.full((Q,), -1, dtype=torch.int64, device=dev)
        pert = synth.make_device(n_planted, D, seed_noise=synth.SEED_QUERY + 1, device=dev)[0]  # plain u in [-1, 1)
        for b_ in range(n_blk):
            q0, q1 = b_ * blk, min((b_ + 1) * blk, n_planted)
            r0 = (b_ * (n_total - blk)) // max(n_blk - 1, 1) if n_total > blk else 0   # blocks spread over the bank
            rows = synth.make_device(q1 - q0, D, row_offset=r0, n_classes=KNN_CLASSES, noise=0.5, device=dev)[0]
            queries[q0:q1] = rows + 0.25 * pert[q0:q1]
            planted[q0:q1] = torch.arange(r0, r0 + (q1 - q0), device=dev) % max(n_total, 1)
        del pert
        kw, kk = 3, max(1, args.knn_steps)
        for _ in range(kw):
            clf.kneighbors_device(queries)
        barrier()
        launch_count_reset()
        kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(kk)]
        for i in range(kk):
            kev[i][0].record()
            dist_d, ids_d = clf.kneighbors_device(queries)
            kev[i][1].record()
        barrier()
        knn_launches = launch_count()
        knn_uncertified = clf.last_uncertified
        # parity gate inside the bench: top-1 of every planted query is the planted row, on every rank; the id table
        # is hashed so that runs at N = 1, 2, 4, 8 can be compared (the search is exact: the hash must not change)
        import hashlib

        ok_planted = bool((ids_d[:n_planted, 0] == planted[:n_planted]).all().item()) if n_total >= blk else None
        all_ok = max_over_ranks(0.0 if ok_planted in (True, None) else 1.0) == 0.0
        ids_sha = hashlib.sha256(ids_d.cpu().numpy().tobytes()).hexdigest()
        if not all_ok:
            raise SystemExit("bench.py: kNN parity gate failed: a planted query's nearest row is not the planted row")
        knn_ms = max_over_ranks(sum(a.elapsed_time(b) for a, b in kev))
        knn_value = Q * kk / (knn_ms * 1e-3)
        lib.en_prof_enable(1)
        clf.kneighbors_device(queries)
        ms = ctypes.c_float(0)
        _lib.check(lib.en_prof_last_ms(ctypes.byref(ms)), "en_prof_last_ms")
        lib.en_prof_enable(0)
        scan_ms = max_over_ranks(ms.value)
        kflops = 2.0 * Q * (hi - lo) * D
        kach = kflops / (scan_ms * 1e-3) / 1e12
        kpeak = peaks["bf16_sustained"] / 3.0  # split-BF16 scan: 3 kind::f16 MMAs per k-step
        # end to end: pinned host queries in, (dist, ids) out
        q_h = queries.cpu().pin_memory()
        out_d = torch.empty((Q, KNN_K), dtype=torch.float32).pin_memory()
        out_i = torch.empty((Q, KNN_K), dtype=torch.int64).pin_memory()
        barrier()
        t0 = time.perf_counter()
        qd = q_h.to(dev, non_blocking=True)
        dd, ii = clf.kneighbors_device(qd)
        out_d.copy_(dd, non_blocking=True)
        out_i.copy_(ii, non_blocking=True)
        torch.cuda.synchronize()
        knn_e2e_dt = max_over_ranks(time.perf_counter() - t0)
        # HBM-bound variant: the reference's one-image-per-call pattern (models.py:122,135), and 8 queries per pass
        stream_bytes = float(hi - lo) * D * 4
        stream = {}
        for qn in (1, 8):
            qs_ = queries[:qn].contiguous()
            for _ in range(3):
                clf.kneighbors_device(qs_)
            lib.en_prof_enable(1)
            sms = []
            for _ in range(5):
                clf.kneighbors_device(qs_)
                ms = ctypes.c_float(0)
                _lib.check(lib.en_prof_last_ms(ctypes.byref(ms)), "en_prof_last_ms")
                sms.append(ms.value)
            lib.en_prof_enable(0)
            k_ms = max_over_ranks(sum(sms) / len(sms))
            barrier()
            s0 = torch.cuda.Event(enable_timing=True)
            s1 = torch.cuda.Event(enable_timing=True)
            s0.record()
            for _ in range(5):
                clf.kneighbors_device(qs_)
            s1.record()
            barrier()
            call_ms = max_over_ranks(s0.elapsed_time(s1) / 5)
            on_stream = qn <= clf.stream_max_q
            stream[qn] = {
                "what": ("%d query(ies) per call, CUDA-core fp32 streaming scan of the bank shard + exact re-rank" if
                         on_stream else "%d query(ies) per call, bank-stationary tcgen05 scan of the shard's BF16 planes "
                         "(same bytes as the fp32 rows; queries resident in shared memory) + exact re-rank") % qn,
                "queries_per_sec": qn / (call_ms * 1e-3), "ms_per_call": call_ms,
                "roofline": {"bound": "hbm", "achieved": stream_bytes / (k_ms * 1e-3) / 1e9, "peak": peaks["hbm_gbs"],
                             "unit": "GB/s", "frac": stream_bytes / (k_ms * 1e-3) / 1e9 / peaks["hbm_gbs"],
                             "traffic": ncu_traffic("knn_stream_q%d_dram_bytes_per_launch" % qn),
                             "kernel": ("knn_stream_kernel<8,4,%d,true>" % qn) if on_stream else
                             "knn_smallq_kernel<16> (bank = 128-row MMA operand)", "kernel_ms": k_ms,
                             "algorithmic_bytes_per_launch": stream_bytes}}
        # C4: offline hard-negative mining over a 1M x 256 bank with the generator's strategies.  Two banks: the
        # SURVEY 8(d) one (noise 0.5: classes well separated, few pairs have a semi-hard candidate) and a candidate-rich
        # one (noise 0.6: classes overlap, most pairs have candidates); 64k-anchor subset for both, plus ALL 1M rows as
        # anchors for the hardest strategy ("anchors = all rows (report also a 64k-anchor subset)").
        mining = None
        if args.knn_bank >= 1_000_000:
            del clf, bank, queries
            torch.cuda.empty_cache()
            n4 = 1_000_000
            lo4, hi4 = BankKNNClassifier.shard_bounds(n4, world, rank)
            ids4 = (torch.arange(n4, dtype=torch.int64, device=dev) % 10_000).to(torch.int32)
            a_idx = torch.arange(0, n4, n4 // 65536, device=dev)[:65536]
            a_lab = ids4[a_idx].contiguous()

            def mining_bank(noise):
                full4 = lac.l2_normalize(synth.make_device(n4, 256, n_classes=10_000, noise=noise, device=dev)[0]).detach()
                bank4 = full4[lo4:hi4].contiguous()   # embeddings_normalization=True, the reference default
                clf4 = BankKNNClassifier(n_neighbors=1, process_group=group, device=dev)
                clf4.fit_shard(bank4, ids4, lo4, n4, classes=np.arange(10_000))
                anchors = full4[a_idx].contiguous()
                positives = full4[(a_idx + 10_000) % n4].unsqueeze(1).contiguous()  # label = id % 10000: same class
                return full4, clf4, anchors, positives

            def timed_semihard(clf4, anchors, positives, what):
                np.random.seed(0)
                clf4.mine_negatives(anchors[:4096], a_lab[:4096], positives=positives[:4096], margin=MARGIN,
                                    mode="semihard")
                barrier()
                t0 = time.perf_counter()
                sel = clf4.mine_negatives(anchors, a_lab, positives=positives, margin=MARGIN, mode="semihard")
                torch.cuda.synchronize()
                s_dt = max_over_ranks(time.perf_counter() - t0)
                return {"workload": "C4: semihard negative (datagenerators.py:196-199 over the whole bank) for 65536 "
                                    "(anchor, positive) pairs, 1M x 256 bank (%s), %d GPU(s); count scan + host RNG "
                                    "draws + select scan over the anchors that own a draw" % (what, world),
                        "ms": s_dt * 1e3, "pairs_per_sec": 65536 / s_dt, "pairs_with_a_candidate": int((sel >= 0).sum())}

            full4, clf4, anchors, positives = mining_bank(0.5)
            for _ in range(2):
                clf4.kneighbors_device(anchors, n_neighbors=1, exclude_labels=a_lab)
            barrier()
            m0, m1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            m0.record()
            clf4.kneighbors_device(anchors, n_neighbors=1, exclude_labels=a_lab)
            m1.record()
            barrier()
            m_ms = max_over_ranks(m0.elapsed_time(m1))
            mining = {"workload": "C4: hardest negative of 65536 anchors over a 1M x 256 bank (label-excluded 1-NN), "
                                  "%d GPU(s)" % world, "ms": m_ms, "anchors_per_sec": 65536 / (m_ms * 1e-3)}
            mining["semihard"] = timed_semihard(clf4, anchors, positives, "noise 0.5: separated classes")
            # all 1M rows as anchors (hardest strategy): 2 * (1e6)^2 * 256 = 512 TFLOP algorithmic
            barrier()
            f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            f0.record()
            clf4.kneighbors_device(full4, n_neighbors=1, exclude_labels=ids4)
            f1.record()
            barrier()
            f_ms = max_over_ranks(f0.elapsed_time(f1))
            mining["all_rows_hardest"] = {
                "workload": "C4 at its stated size: hardest negative of ALL 1,000,000 rows over the 1M x 256 bank, "
                            "%d GPU(s)" % world, "ms": f_ms, "anchors_per_sec": n4 / (f_ms * 1e-3),
                "tflops_algorithmic": 2.0 * n4 * n4 * 256 / (f_ms * 1e-3) / 1e12,
                "frac_of_bf16x3_roofline": 2.0 * n4 * n4 * 256 / (f_ms * 1e-3) / 1e12 / world /
                                           (peaks["bf16_sustained"] / 3.0)}
            del full4, clf4, anchors, positives
            torch.cuda.empty_cache()
            full4, clf4, anchors, positives = mining_bank(0.6)
            mining["semihard_candidate_rich"] = timed_semihard(clf4, anchors, positives,
                                                               "noise 0.6: overlapping classes")
            del full4, clf4, anchors, positives
        # Key order matters to the reader of a truncated log: bulky sub-records first, the headline of the sharded
        # path (value, time, roofline, e2e, parity hash, NCCL ranks) last.
        knn = {
            "metric": "knn_queries_per_sec_10M_bank", "unit": "queries/s", "scaling": "strong",
            "config": {"workload": "C5: %d queries (90%% planted: bank row + 0.25 u, 10%% unrelated) vs %d x %d fp32 "
                       "bank, k=%d, bank sharded row-wise over %d GPU(s), NCCL all-gather + merge; inputs >> L2 (no "
                       "flush needed)" % (Q, n_total, D, KNN_K, world)},
            "stream_scan": stream[1], "stream_scan_q8": stream[8],
        }
        if mining is not None:
            knn["bank_mining"] = mining
        if rank == 0 and world == 1 and not args.skip_cpu:
            qps, sample = cpu_knn_baseline()
            knn["cpu_baseline"] = {"value": qps, "unit": "queries/s", "cores": os.cpu_count(), "kind": "port",
                                   "sample": sample}
        knn.update({
            "gpu_launches": int(knn_launches),
            "certificate": {"uncertified_queries_this_rank": int(knn_uncertified), "of": Q,
                            "note": "queries whose exactness proof failed are redone by float64 brute force "
                                    "inside the timed call"},
            "roofline": {"bound": "tensor", "achieved": kach, "peak": kpeak, "unit": "TFLOP/s", "frac": kach / kpeak,
                         "traffic": ncu_traffic("knn_scan_dram_bytes_per_launch"),
                         "kernel": "dist_gemm_kernel<EpTopK<8>> (split-BF16 planes)",
                         "kernel_ms": scan_ms, "share_of_step": scan_ms / (knn_ms / kk),
                         "algorithmic_flops_per_launch": kflops,
                         "peak_note": "%s bf16 dense %.1f TFLOP/s (sustained) / 3, per GPU" % (
                             peaks["source"], peaks["bf16_sustained"]),
                         "frac_of_tf32x3_roofline": kach / (peaks["bf16_sustained"] / 6.0)},
            "e2e": {"value": Q / knn_e2e_dt, "unit": "queries/s", "h2d_bytes_per_step": Q * D * 4,
                    "d2h_bytes_per_step": Q * KNN_K * 12, "api": "BankKNNClassifier.kneighbors (pinned host queries)"},
            "parity": {"planted_top1_ok": ok_planted, "planted_queries": int(n_planted), "ids_sha256": ids_sha},
            "n_gpus": world, "nccl_ranks": (dist.get_world_size() if world > 1 else 1),
            "steps": kk, "warmup": kw, "ms_per_step": knn_ms / kk, "value": knn_value,
        })
        line["knn"] = knn

    clocks = sampler.stop()
    # `knn` stays the last key: the tail of the line is the sharded path's headline
    tail = line.pop("knn", None)
    line["clocks"] = clocks
    if tail is not None:
        line["knn"] = tail
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
        if nccl_log and os.path.exists(nccl_log):
            with open(nccl_log) as f:
                keep = [ln for ln in f if "nranks" in ln or "Init COMPLETE" in ln or "NCCL version" in ln]
            sys.stderr.write("".join(keep[:8]))
            sys.stderr.flush()
            os.remove(nccl_log)


if __name__ == "__main__":
    main()
