#!/usr/bin/env python
"""bench.py -- GCUPS (incl. traceback) of the batched align hot path on 1..8 B200.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...   # CPU restatement of the Go path

A "step" is one pass of the hot path (encode/validate, DP fill, traceback walk, segment
gather) over one shard of synthetic pairs of the BASELINE config (default config 2:
align.SWAffine, 250 bp reads vs 1 kb windows; 125 000 pairs per GPU = 1 M pairs over 8 GPUs,
weak scaling).  Pairs are independent, so ranks share nothing: no collective on the data
path, only a barrier + max-over-ranks for timing.

value   device-timed (CUDA events on the launching stream), inputs resident in HBM
e2e     the same metric through ba_align_batch with HOST buffers (pinned), H2D + D2H inside
roofline  traceback bytes written by the fill kernels / fill-kernel time vs measured HBM peak,
          plus the DPX/integer view (lane-results per cell) under "dpx"
cpu_baseline  the C oracle (port of the Go path) on all host threads, bounded sample
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "GCUPS incl. traceback, batched SWAffine/NWAffine/FittedAffine at 1/2/4/8 B200"
DEFAULT_PAIRS = {1: 100_000, 2: 125_000, 3: 125_000, 4: 16, 5: 25_000}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d.get("hbm_gbs", 6650.0)), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def load_dpx_peak():
    """lane-results / clk / SM for packed 16-bit DPX, from profiles/ (our micro-benchmark)."""
    p = os.path.join(ROOT, "profiles", "dpx_peaks.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f)
    return None


class ClockSampler:
    """SM clocks + throttle reasons sampled DURING the timed region: NVML from a thread (a few ms
    per sample, so even a 100 ms region gets tens of samples); nvidia-smi -lms as the fallback."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int, uuid: str | None = None):
        self.idx, self.proc, self.lines = gpu_index, None, []
        self.nvml, self.h, self.samples, self.run = None, None, [], False
        try:
            import pynvml

            pynvml.nvmlInit()
            h = None
            if uuid:
                try:
                    h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
                except Exception:  # noqa: BLE001
                    h = None
            self.h = h or pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
            self.nvml = pynvml
        except Exception:  # noqa: BLE001
            self.nvml = None

    def _nvml_loop(self):
        n = self.nvml
        while self.run:
            try:
                sm = n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM)
                mx = n.nvmlDeviceGetMaxClockInfo(self.h, n.NVML_CLOCK_SM)
                rs = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.samples.append((sm, mx, rs))
            except Exception:  # noqa: BLE001
                break
            time.sleep(0.004)

    def _nvml_stop(self):
        self.run = False
        self.t.join(timeout=2)
        n = self.nvml
        if not self.samples:
            return None
        sm = sorted(x[0] for x in self.samples)
        reasons = set()
        names = (("hw_slowdown", getattr(n, "nvmlClocksThrottleReasonHwSlowdown", 0x8)),
                 ("hw_thermal_slowdown", getattr(n, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40)),
                 ("sw_thermal_slowdown", getattr(n, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20)),
                 ("sw_power_cap", getattr(n, "nvmlClocksThrottleReasonSwPowerCap", 0x4)))
        for _, _, rs in self.samples:
            for name, bit in names:
                if rs & bit:
                    reasons.add(name)
        return {"sm_mhz": float(sm[len(sm) // 2]), "sm_max_mhz": float(max(x[1] for x in self.samples)),
                "reasons": sorted(reasons), "samples": len(self.samples), "source": "nvml"}

    def start(self):
        if self.nvml:
            self.run = True
            self.t = threading.Thread(target=self._nvml_loop, daemon=True)
            self.t.start()
            return
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.nvml:
            r = self._nvml_stop()
            if r:
                return r
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        # keep only samples under load (clock above idle)
        load = [c for c in sm if c > 0.5 * (max(mx) if mx else 1)] or sm
        return {"sm_mhz": float(np.median(load)) if load else None,
                "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


def cpu_port_gcups(w, sample_pairs: int, threads: int, reps: int = 1):
    """The oracle (C port of the Go path: full int64 table per pair) on `threads` host threads."""
    from oracle import oracle as O

    n = min(sample_pairs, w.n)
    sub = w.subset(np.arange(n))
    sc = O.Scoring(w.kind, w.affine, w.matrix, w.gap_open if w.affine else 0, w.alpha.LetterIndex())
    best = None
    for _ in range(reps):
        t0 = time.perf_counter()
        O.align_batch(sc, sub.letters, sub.stride, sub.ref_off, sub.ref_len, sub.qry_off, sub.qry_len, threads)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return sub.cells / best / 1e9, n, sub.cells, best


def workload_config(name: str, config_num: int, pairs_per_gpu: int, world: int) -> dict:
    """The `config` object of the JSON line -- the SAME for the GPU arm and the reference arm: it names the
    workload (BASELINE.json config, shape, per-GPU batch of the weak-scaling job), not what a particular
    arm sampled from it (that is in `cpu_baseline.sample` / `config_sample`)."""
    return {"workload": name, "baseline_config": config_num, "pairs_per_gpu_per_step": pairs_per_gpu,
            "parallelism": f"independent pairs sharded over {world} GPU(s), no collective",
            "l2": "inputs + traceback checkpoints per step exceed the 126 MB L2",
            "seed": "splitmix64 0xB10600+config (+rank<<20)"}


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path.  Go cannot run in
    this image, so it is the oracle port (oracle/align_oracle.c), all host threads."""
    if rank != 0:
        return
    from biogo_b200 import synth

    threads = os.cpu_count() or 1
    per_step = args.cpu_sample or max(64, 1000 * threads)
    w = synth.config(args.config, per_step)
    for _ in range(args.warmup):
        cpu_port_gcups(w, min(per_step, 4 * threads), threads)
    t0 = time.perf_counter()
    cells = 0
    for _ in range(args.steps):
        g, n, c, dt = cpu_port_gcups(w, per_step, threads)
        cells += c
    dt = time.perf_counter() - t0
    val = cells / dt / 1e9
    pairs = args.pairs_per_gpu or DEFAULT_PAIRS[args.config]
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "GCUPS", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int64",
        "data": "synthetic",
        "config": workload_config(w.name, args.config, pairs, world),
        "config_sample": {"pairs_per_step": w.n, "cells_per_step": w.cells,
                          "note": "bounded sample of the same workload (same generator, same seed): the first "
                                  f"{w.n} pairs of the step, so the run ends within minutes"},
        "cpu_baseline": {"value": val, "unit": "GCUPS", "cores": threads, "kind": "port",
                         "sample": f"{w.n} pairs of {w.name} per step, {threads} pthreads"},
        "e2e": {"value": val, "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


_REAL_STDOUT = None


def emit(line: dict) -> None:
    """Write the one JSON line to the process's original stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def bind_near_gpu(local_rank: int):
    """Multi-rank runs: keep this rank's host threads (and therefore its pinned result / staging buffers, which
    are placed by first touch) on the CPUs NVML reports as local to its GPU.  Eight ranks copying 16 MB of
    segments per step otherwise cross sockets at random.  Returns the CPU list, or None when unavailable."""
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = [64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1]
        cpus = [c for c in cpus if c in os.sched_getaffinity(0)]
        if cpus:
            os.sched_setaffinity(0, cpus)
            return cpus
    except Exception:  # noqa: BLE001
        pass
    return None


ALG_BYTES_PER_CELL = {True: 1.0, False: 0.25}   # affine 1 byte/cell, linear 2 bits/cell
AGG_KEYS = ("ms_fill", "ms_trace", "ms_host", "ms_kernels", "ms_d2h", "gpu_launches", "fill_launches", "code_bytes",
            "n_segs", "pairs_fast16", "waves")


class DeviceBatch:
    """A workload resident in HBM + a context with its scoring set: one `step()` = one pass of the hot path."""

    def __init__(self, torch, _lib, w, dev, local_rank):
        self.w, self.torch = w, torch
        self.ctx = _lib.Context(local_rank, os.environ.get("BA_LIB_PATH") or None)
        self.ctx.set_scoring(w.kind, w.affine, np.array(w.matrix, np.int64).reshape(-1), len(w.matrix), w.alpha.Len(),
                             w.gap_open if w.affine else 0, w.alpha.LetterIndex())
        self.t = [torch.from_numpy(a).to(dev) for a in (w.letters, w.ref_off, w.ref_len, w.qry_off, w.qry_len)]

    def step(self, stream=0):
        t, w = self.t, self.w
        return self.ctx.align_device(t[0].data_ptr(), t[0].numel(), w.stride, t[1].data_ptr(), t[2].data_ptr(),
                                     t[3].data_ptr(), t[4].data_ptr(), w.ref_len, w.qry_len, stream=stream,
                                     collect=False)


def timed_steps(torch, dev, batch, steps, warmup, stream):
    """W untimed + K timed steps, CUDA events on the launching stream; returns (ms, aggregated stats)."""
    for _ in range(max(warmup, 0)):
        batch.step(stream.cuda_stream)
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    agg = {k: 0.0 for k in AGG_KEYS}
    e0.record(stream)
    for _ in range(steps):
        s = batch.step(stream.cuda_stream)
        for k in agg:
            agg[k] += s[k]
    e1.record(stream)
    torch.cuda.synchronize(dev)
    return e0.elapsed_time(e1), agg


def roofline_of(w, cells, steps, agg, hbm_peak):
    """Algorithmic-byte roofline of the fill kernels: bytes/cell x cells / fill time vs measured HBM."""
    fill_s = agg["ms_fill"] * 1e-3
    alg = ALG_BYTES_PER_CELL[bool(w.affine)]
    ach = alg * cells * steps / fill_s / 1e9 if fill_s > 0 else 0.0
    return ach, alg, fill_s


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=2, help="BASELINE.json config number (1..5)")
    ap.add_argument("--pairs-per-gpu", type=int, default=0)
    ap.add_argument("--cpu-sample", type=int, default=0, help="pairs in the CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true", help="skip the short runs of the other BASELINE configs")
    ap.add_argument("--no-onebatch", action="store_true", help="skip the one-process multi-GPU batch (ba_align_batch_multi)")
    args = ap.parse_args()

    # Only the JSON line may reach stdout: libraries (NCCL's version banner, torchrun notices) write
    # to file descriptor 1 directly, so park the real stdout and point fd 1 at stderr meanwhile.
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    from biogo_b200 import _lib, synth

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: biogo_b200 has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    all_cpus = os.sched_getaffinity(0)
    affinity = bind_near_gpu(local_rank) if world > 1 else None
    cpu_group = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # keep stdout for the one JSON line
        dist.init_process_group("nccl", device_id=dev)
        cpu_group = dist.new_group(backend="gloo")   # host-side waits that must not occupy the GPUs

    pairs = args.pairs_per_gpu or DEFAULT_PAIRS[args.config]
    w = synth.config(args.config, pairs, shard=rank)
    batch = DeviceBatch(torch, _lib, w, dev, local_rank)
    ctx = batch.ctx
    stream = torch.cuda.current_stream(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(max(args.warmup, 0)):
        batch.step(stream.cuda_stream)
    try:
        gpu_uuid = str(torch.cuda.get_device_properties(dev).uuid)
    except Exception:  # noqa: BLE001
        gpu_uuid = None
    sampler = ClockSampler(local_rank, gpu_uuid)
    barrier()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    agg = {k: 0.0 for k in AGG_KEYS}
    for _ in range(args.steps):
        s = batch.step(stream.cuda_stream)
        for k in agg:
            agg[k] += s[k]
    e1.record(stream)
    barrier()
    clocks = sampler.stop()
    ms = e0.elapsed_time(e1)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    cells_step = w.cells
    value = world * cells_step * args.steps / (ms_max * 1e-3) / 1e9

    # ---- the same steps with the plan cache off (every call plans, sorts and uploads its work lists: what a
    # stream of batches with varying read lengths pays)
    plan_miss = None
    try:
        ctx.set_option("plan_cache", 0)
        k = max(2, min(args.steps, 5))
        ms_miss, agg_miss = timed_steps(torch, dev, batch, k, 1, stream)
        plan_miss = {"ms_per_step": ms_miss / k, "value": cells_step / (ms_miss / k * 1e-3) / 1e9,
                     "host_plan_ms_per_step": agg_miss["ms_host"] / k, "steps": k}
    except Exception as ex:  # noqa: BLE001
        plan_miss = {"error": str(ex)[:200]}
    finally:
        ctx.set_option("plan_cache", 1)
        batch.step(stream.cuda_stream)

    # ---- end to end from pinned host buffers (for `e2e`)
    e2e = None
    if not args.no_e2e:
        import ctypes as C

        nb = w.letters.nbytes
        hp = ctx.lib.ba_alloc_host(max(nb, 1))
        if not hp:
            raise SystemExit("ba_alloc_host failed")
        h_letters = np.ctypeslib.as_array(C.cast(hp, C.POINTER(C.c_uint8)), shape=(max(nb, 1),))[:nb]
        h_letters[:] = w.letters

        def e2e_step(cx):
            return cx.align_packed(h_letters, w.stride, w.ref_off, w.ref_len, w.qry_off, w.qry_len, collect=False)

        def wall(fn):
            barrier()
            t0 = time.perf_counter()
            fn()
            torch.cuda.synchronize(dev)
            tt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            return float(tt.item())

        # (1) one call at a time: the upload travels in groups and the fill starts on the first of them
        e2e_step(ctx)
        e2e_step(ctx)
        segs = [0]

        def serial():
            for _ in range(args.steps):
                segs[0] = e2e_step(ctx)["n_segs"]
        dt_serial = wall(serial)

        # (2) the way bíogo's aligners are driven (many goroutines, concurrent/processor.go:48-74):
        # two host threads, each with its own ba_ctx, take alternate batches, so the copies and the tail of
        # one batch overlap the kernels of the other.  Every step still pays its own copies.
        dt = dt_serial
        value_two = None
        mode = "one call at a time (the overlapped run failed: %s)"
        try:
            ctx2 = _lib.Context(local_rank, os.environ.get("BA_LIB_PATH") or None)
            ctx2.set_scoring(w.kind, w.affine, np.array(w.matrix, np.int64).reshape(-1), len(w.matrix), w.alpha.Len(),
                             w.gap_open if w.affine else 0, w.alpha.LetterIndex())
            e2e_step(ctx2)
            errs = []

            def worker(cx, k):
                try:
                    for _ in range(k):
                        e2e_step(cx)
                except Exception as ex:  # noqa: BLE001
                    errs.append(ex)

            def overlapped():
                th = [threading.Thread(target=worker, args=(ctx, (args.steps + 1) // 2)),
                      threading.Thread(target=worker, args=(ctx2, args.steps // 2))]
                for x in th:
                    x.start()
                for x in th:
                    x.join()
            dt = wall(overlapped)
            if errs:
                raise errs[0]

            # the same two-context scheme on the HBM-resident inputs (each context on its own stream)
            t_ = batch.t

            def dev_worker(cx, k):
                try:
                    for _ in range(k):
                        cx.align_device(t_[0].data_ptr(), t_[0].numel(), w.stride, t_[1].data_ptr(), t_[2].data_ptr(),
                                        t_[3].data_ptr(), t_[4].data_ptr(), w.ref_len, w.qry_len, stream=0, collect=False)
                except Exception as ex:  # noqa: BLE001
                    errs.append(ex)

            def dev_overlapped():
                th = [threading.Thread(target=dev_worker, args=(ctx, (args.steps + 1) // 2)),
                      threading.Thread(target=dev_worker, args=(ctx2, args.steps // 2))]
                for x in th:
                    x.start()
                for x in th:
                    x.join()
            dev_worker(ctx2, 1)
            dt_dev2 = wall(dev_overlapped)
            if errs:
                raise errs[0]
            value_two = world * cells_step * args.steps / dt_dev2 / 1e9
            ctx2.close()
            del ctx2
            mode = None
        except Exception as ex:  # noqa: BLE001 -- keep the single-call number rather than lose the line
            dt = dt_serial
            mode = mode % (str(ex)[:120],)
        h2d = nb + w.n * (8 + 8 + 4 + 4)
        d2h = segs[0] * 20 + w.n * (4 + 8 + 8 + 4)
        e2e = {"value": world * cells_step * args.steps / dt / 1e9, "unit": "GCUPS",
               "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "single_call_value": world * cells_step * args.steps / dt_serial / 1e9,
               "device_resident_two_contexts_value": value_two,
               "timing": "host wall clock around ba_align_batch calls on pinned host buffers, max over ranks; "
                         "value: " + (mode or "2 host threads with one ba_ctx each take alternate batches (copies and "
                                              "tail of one overlap kernels of the other)") +
                         "; single_call_value: one blocking call at a time (its upload travels in groups, the fill "
                         "starts chunk by chunk on what has arrived); device_resident_two_contexts_value: the "
                         "two-thread scheme on HBM-resident inputs.  The overlapped numbers can exceed the top-level "
                         "`value`, which runs one context step after step (CUDA-event timed): there the result "
                         "copy and host work of a step are not hidden under the next step's kernels"}
        ctx.lib.ba_free_host(hp)

    if affinity:
        os.sched_setaffinity(0, all_cpus)   # the one-process multi-GPU batch and the CPU baseline use every core
    # ---- the other BASELINE configs, a few steps each on this rank's GPU (rank 0 reports them)
    other = {}
    hbm_peak, peak_src = load_peaks()
    if not args.no_other_configs and rank == 0:
        for num, n_pairs in ((1, DEFAULT_PAIRS[1]), (3, DEFAULT_PAIRS[3]), (4, 128), (5, DEFAULT_PAIRS[5])):
            if num == args.config:
                continue
            try:
                wo = synth.config(num, n_pairs)
                bo = DeviceBatch(torch, _lib, wo, dev, local_rank)
                k = 4
                # median of three rounds of k steps (the first round after a fresh context still pays allocations)
                rounds = sorted((timed_steps(torch, dev, bo, k, 3, stream) for _ in range(3)), key=lambda r: r[0])
                ms_o, agg_o = rounds[1]
                ach, alg, fill_s = roofline_of(wo, wo.cells, k, agg_o, hbm_peak)
                other[f"cfg{num}"] = {
                    "workload": wo.name, "pairs": wo.n, "steps": k, "warmup": 3, "rounds": "median of 3",
                    "value": wo.cells * k / (ms_o * 1e-3) / 1e9, "unit": "GCUPS", "ms_per_step": ms_o / k,
                    "fill_ms_per_step": agg_o["ms_fill"] / k, "trace_ms_per_step": agg_o["ms_trace"] / k,
                    "roofline_frac_algorithmic": ach / hbm_peak, "algorithmic_bytes_per_cell": alg,
                    "stored_bytes_per_cell": agg_o["code_bytes"] / (wo.cells * k),
                    "dtype": "int16" if agg_o["pairs_fast16"] else "int32",
                    "pairs_on_dpx16_path": int(agg_o["pairs_fast16"] // k)}
                bo.ctx.close()
                del bo, wo
                torch.cuda.empty_cache()
            except Exception as ex:  # noqa: BLE001
                other[f"cfg{num}"] = {"error": str(ex)[:200]}

    # ---- ONE batch over all N GPUs from ONE process (ba_create_multi / ba_align_batch_multi): rank 0 drives
    # GPUs 0..N-1 while the other ranks wait on the host (gloo), strong scaling of a fixed batch
    onebatch = None
    if not args.no_onebatch:
        if world > 1:
            ctx.close()                       # the per-rank contexts give their arenas back first
            batch.t = None
            torch.cuda.empty_cache()
            dist.barrier(group=cpu_group)
        if rank == 0:
            onebatch = run_onebatch(_lib, synth, w, args, world)
        if world > 1:
            dist.barrier(group=cpu_group)

    if rank == 0:
        fill_s = agg["ms_fill"] * 1e-3
        ach, alg, _ = roofline_of(w, cells_step, args.steps, agg, hbm_peak)
        traffic, traffic_src, ent = None, None, None
        try:  # DRAM bytes per launch of the fill kernel, captured once with ncu --set full
            with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
                ent = json.load(f)["entries"].get(w.name)
            if ent and ent["pairs"] == w.n and agg["pairs_fast16"]:
                traffic = ent["dram_bytes_read"] + ent["dram_bytes_write"]
                traffic_src = "profiles/ncu_traffic.json (" + ent["kernel"] + ", one launch)"
        except (OSError, KeyError, ValueError):
            pass
        launches = max(agg["fill_launches"], 1)
        alg_per_launch = alg * cells_step * args.steps / launches
        roof = {"bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s",
                "frac": ach / hbm_peak, "traffic": traffic, "traffic_source": traffic_src,
                "traffic_ratio": (traffic / alg_per_launch) if traffic else None,
                "algorithmic_bytes_per_launch": alg_per_launch,
                "peak_source": peak_src,
                "kernel": ("ba_fill16_kernel (packed s16x2 DPX fill + checkpoint stores)" if agg["pairs_fast16"]
                           else "ba_fill32_kernel (DP fill + traceback-code stores)"),
                "algorithmic_bytes_per_cell": alg,
                "stored_bytes_per_cell": agg["code_bytes"] / (cells_step * args.steps),
                "stored_bytes_per_launch": agg["code_bytes"] / launches,
                "fill_ms_per_step": agg["ms_fill"] / args.steps,
                "fill_share_of_step": agg["ms_fill"] / ms if ms > 0 else None,
                "fill_gcups": cells_step * args.steps / fill_s / 1e9 if fill_s > 0 else None,
                "note": "achieved = SURVEY 8(d) algorithmic bytes per cell x cells / fill-kernel time (CUDA events "
                        "inside the step); the checkpoints actually stored are stored_bytes_per_cell"}
        dpx = load_dpx_peak()
        if dpx and fill_s > 0:
            lanes_per_cell = 3 if w.affine else 1
            sm_mhz = clocks.get("sm_mhz") or dpx.get("sm_mhz", 1965.0)
            peak_lane = dpx["s16x2_lane_results_per_clk_per_sm"] * dpx.get("sms", 148) * sm_mhz * 1e6
            achieved_lane = cells_step * args.steps * lanes_per_cell / fill_s
            roof["dpx"] = {"achieved_lane_results_per_s": achieved_lane, "peak_lane_results_per_s": peak_lane,
                           "frac": achieved_lane / peak_lane, "lane_results_per_cell": lanes_per_cell,
                           "peak_source": "profiles/dpx_peaks.json (ba_microbench on B200)"}
            if traffic is not None and ent:  # the same ncu capture: how busy the integer/DPX pipe really is
                roof["dpx"]["ncu_alu_pipe_pct"] = ent.get("alu_pipe_pct")
                roof["dpx"]["ncu_issue_active_pct"] = ent.get("issue_active_pct")
                roof["dpx"]["ncu_thread_inst_per_cell"] = ent.get("inst_per_cell")
        cfg = workload_config(w.name, args.config, w.n, world)   # identical in the reference arm's line
        cfg_detail = {"cells_per_gpu_per_step": cells_step,
                      "host_affinity": (f"each rank bound to the {len(affinity)} CPUs NVML reports as local to its GPU"
                                        if affinity else "not bound"),
                      "host_plan": "work lists cached across same-shaped batches (lengths only, never results); "
                                   "the uncached cost is plan_cache_miss"}
        line = {
            "metric": METRIC, "value": value, "unit": "GCUPS", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int16" if agg["pairs_fast16"] else "int32",
            "data": "synthetic", "config": cfg, "config_detail": cfg_detail,
            "roofline": roof, "clocks": clocks,
            "breakdown_ms_per_step": {"fill": agg["ms_fill"] / args.steps, "trace": agg["ms_trace"] / args.steps,
                                      "d2h": agg["ms_d2h"] / args.steps, "host_plan_scatter": agg["ms_host"] / args.steps,
                                      "kernels_span": agg["ms_kernels"] / args.steps,
                                      "waves": agg["waves"] / args.steps,
                                      "pairs_on_dpx16_path": int(agg["pairs_fast16"] // args.steps),
                                      "note": "CUDA-event spans inside the library; per-pair results are gathered in "
                                              "pair order on the device, the host neither scans nor scatters"},
            "plan_cache_miss": plan_miss,
            "gpu_launches": int(agg["gpu_launches"]),
        }
        if e2e:
            line["e2e"] = e2e
        if other:
            line["other_configs"] = other
        if onebatch:
            line["onebatch"] = onebatch
        if not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            sample = args.cpu_sample or max(64, 2000 * threads)
            g, n, c, dtc = cpu_port_gcups(w, sample, threads)
            line["cpu_baseline"] = {"value": g, "unit": "GCUPS", "cores": threads, "kind": "port",
                                    "sample": f"first {n} pairs of the step ({c} cells), C port of the Go path "
                                              f"(full int64 table per pair), {threads} pthreads, {dtc:.1f} s"}
        emit(line)
    if world > 1:
        dist.barrier(group=cpu_group)
        dist.destroy_process_group()


def run_onebatch(_lib, synth, w, args, world):
    """north_star's multi-GPU path: ONE batch, ONE process, N GPUs, host buffers in, gathered results out
    (ba_align_batch_multi: dealing + per-device gather/upload + kernels + result gather, all inside the timed
    region).  Strong scaling: the batch is the same for every N (8 x the per-GPU shard of the headline config,
    i.e. 1 M pairs of config 2) and, second, 8 x the 25 000-pair shard of config 5 (mixed lengths, QLetters:
    dealt by length buckets and gathered)."""
    import ctypes as C

    out = {"n_gpus": world, "process": "one process, one host thread per device, no collective",
           "timing": "host wall clock around blocking ba_align_batch_multi calls on pinned host buffers",
           "batches": {}}
    try:
        mc = _lib.MultiContext(tuple(range(world)), os.environ.get("BA_LIB_PATH") or None)
    except Exception as ex:  # noqa: BLE001
        return {"error": str(ex)[:200]}
    todo = [("cfg%d" % args.config, w, 8)]
    if args.config != 5:
        try:
            todo.append(("cfg5", synth.config(5, DEFAULT_PAIRS[5]), 8))
        except Exception as ex:  # noqa: BLE001
            out["batches"]["cfg5"] = {"error": str(ex)[:200]}
    for name, ww, reps in todo:
        hp = None
        try:
            nb = ww.letters.nbytes
            hp = mc.lib.ba_alloc_host(nb * reps)
            if not hp:
                raise RuntimeError("ba_alloc_host failed")
            hl = np.ctypeslib.as_array(C.cast(hp, C.POINTER(C.c_uint8)), shape=(nb * reps,))
            ro, qo = [], []
            for r in range(reps):
                hl[r * nb:(r + 1) * nb] = ww.letters
                ro.append(ww.ref_off + r * nb)
                qo.append(ww.qry_off + r * nb)
            ro, qo = np.concatenate(ro), np.concatenate(qo)
            rl, ql = np.tile(ww.ref_len, reps), np.tile(ww.qry_len, reps)
            cells = ww.cells * reps
            mc.set_scoring(ww.kind, ww.affine, np.array(ww.matrix, np.int64).reshape(-1), len(ww.matrix),
                           ww.alpha.Len(), ww.gap_open if ww.affine else 0, ww.alpha.LetterIndex())
            st = None
            for _ in range(2):
                st = mc.align_packed(hl, ww.stride, ro, rl, qo, ql, collect=False)
            k = 3
            t0 = time.perf_counter()
            for _ in range(k):
                st = mc.align_packed(hl, ww.stride, ro, rl, qo, ql, collect=False)
            dt = (time.perf_counter() - t0) / k
            out["batches"][name] = {
                "workload": ww.name, "pairs": int(len(rl)), "cells": int(cells), "steps": k, "warmup": 2,
                "batch": f"{reps} x the {ww.n}-pair synthetic shard (pairs repeat; lengths and letters as generated)",
                "ms_per_batch": dt * 1e3, "value": cells / dt / 1e9, "unit": "GCUPS",
                "h2d_bytes": int(nb * reps + len(rl) * 24), "d2h_bytes": int(st["n_segs"] * 20 + len(rl) * 24),
                "slowest_device_kernels_span_ms": st["ms_kernels"], "host_deal_ms": st["ms_host"]}
        except Exception as ex:  # noqa: BLE001
            out["batches"][name] = {"error": str(ex)[:200]}
        finally:
            if hp:
                mc.lib.ba_free_host(hp)
    mc.close()
    return out


if __name__ == "__main__":
    main()
