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
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "GCUPS", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int64",
        "data": "synthetic",
        "config": {"workload": w.name, "pairs_per_step": w.n, "cells_per_step": w.cells},
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
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # keep stdout for the one JSON line
        dist.init_process_group("nccl", device_id=dev)

    pairs = args.pairs_per_gpu or DEFAULT_PAIRS[args.config]
    w = synth.config(args.config, pairs, shard=rank)
    ctx = _lib.Context(local_rank, os.environ.get("BA_LIB_PATH") or None)  # override: kernel-variant experiments
    ctx.set_scoring(w.kind, w.affine, np.array(w.matrix, np.int64).reshape(-1), len(w.matrix), w.alpha.Len(),
                    w.gap_open if w.affine else 0, w.alpha.LetterIndex())

    # ---- inputs resident in HBM (for `value`)
    d_letters = torch.from_numpy(w.letters).to(dev)
    d_ro = torch.from_numpy(w.ref_off).to(dev)
    d_rl = torch.from_numpy(w.ref_len).to(dev)
    d_qo = torch.from_numpy(w.qry_off).to(dev)
    d_ql = torch.from_numpy(w.qry_len).to(dev)
    stream = torch.cuda.current_stream(dev)

    def step_device():
        return ctx.align_device(d_letters.data_ptr(), d_letters.numel(), w.stride, d_ro.data_ptr(),
                                d_rl.data_ptr(), d_qo.data_ptr(), d_ql.data_ptr(), w.ref_len, w.qry_len,
                                stream=stream.cuda_stream, collect=False)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(max(args.warmup, 0)):
        step_device()
    try:
        gpu_uuid = str(torch.cuda.get_device_properties(dev).uuid)
    except Exception:  # noqa: BLE001
        gpu_uuid = None
    sampler = ClockSampler(local_rank, gpu_uuid)
    barrier()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    agg = {"ms_fill": 0.0, "ms_trace": 0.0, "ms_host": 0.0, "ms_kernels": 0.0, "ms_d2h": 0.0, "gpu_launches": 0,
           "fill_launches": 0, "code_bytes": 0, "n_segs": 0, "pairs_fast16": 0}
    for _ in range(args.steps):
        s = step_device()
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

        # (1) one call at a time: H2D, kernels and D2H of a batch run back to back
        e2e_step(ctx)
        segs = [0]

        def serial():
            for _ in range(args.steps):
                segs[0] = e2e_step(ctx)["n_segs"]
        dt_serial = wall(serial)

        # (2) the way bíogo's aligners are driven (many goroutines, concurrent/processor.go:48-74):
        # two host threads, each with its own ba_ctx, take alternate batches, so the H2D copy of
        # one batch overlaps the kernels of the other.  Every step still pays its own copies.
        dt = dt_serial
        value_two = None
        mode = "one call at a time (the overlapped run failed: %s)"
        try:
            import threading

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
            def dev_worker(cx, k):
                try:
                    for _ in range(k):
                        cx.align_device(d_letters.data_ptr(), d_letters.numel(), w.stride, d_ro.data_ptr(),
                                        d_rl.data_ptr(), d_qo.data_ptr(), d_ql.data_ptr(), w.ref_len, w.qry_len,
                                        stream=0, collect=False)
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
                         "value: " + (mode or "2 host threads with one ba_ctx each take alternate batches (copies of "
                                              "one overlap kernels of the other)") +
                         "; single_call_value: one call at a time; device_resident_two_contexts_value: the "
                         "two-thread scheme on HBM-resident inputs.  The overlapped numbers can exceed the top-level "
                         "`value`, which runs one context step after step (CUDA-event timed): there the result "
                         "copy and host gather of a step are not hidden under the next step's kernels"}
        ctx.lib.ba_free_host(hp)

    if rank == 0:
        hbm_peak, peak_src = load_peaks()
        fill_s = agg["ms_fill"] * 1e-3
        ach = agg["code_bytes"] / fill_s / 1e9 if fill_s > 0 else 0.0
        traffic, traffic_src = None, None
        try:  # DRAM bytes per launch of the fill kernel, captured once with ncu --set full
            with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
                ent = json.load(f)["entries"].get(w.name)
            if ent and ent["pairs"] == w.n and agg["pairs_fast16"]:
                traffic = ent["dram_bytes_read"] + ent["dram_bytes_write"]
                traffic_src = "profiles/ncu_traffic.json (" + ent["kernel"] + ", one launch)"
        except (OSError, KeyError, ValueError):
            pass
        roof = {"bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s",
                "frac": ach / hbm_peak, "traffic": traffic, "traffic_source": traffic_src,
                "algorithmic_bytes_per_launch": agg["code_bytes"] / max(agg["fill_launches"], 1),
                "peak_source": peak_src,
                "kernel": ("ba_fill16_kernel (packed s16x2 DPX fill + checkpoint stores)" if agg["pairs_fast16"]
                           else "ba_fill32_kernel (DP fill + traceback-code stores)"),
                "algorithmic_bytes_per_cell": (1.0 if w.affine else 0.25),
                "stored_bytes_per_cell": agg["code_bytes"] / (cells_step * args.steps),
                "fill_ms_per_step": agg["ms_fill"] / args.steps,
                "fill_share_of_step": agg["ms_fill"] / ms if ms > 0 else None,
                "fill_gcups": cells_step * args.steps / fill_s / 1e9 if fill_s > 0 else None}
        dpx = load_dpx_peak()
        if dpx and fill_s > 0:
            lanes_per_cell = 3 if w.affine else 1
            sm_mhz = clocks.get("sm_mhz") or dpx.get("sm_mhz", 1965.0)
            peak_lane = dpx["s16x2_lane_results_per_clk_per_sm"] * dpx.get("sms", 148) * sm_mhz * 1e6
            achieved_lane = cells_step * args.steps * lanes_per_cell / fill_s
            roof["dpx"] = {"achieved_lane_results_per_s": achieved_lane, "peak_lane_results_per_s": peak_lane,
                           "frac": achieved_lane / peak_lane, "lane_results_per_cell": lanes_per_cell,
                           "peak_source": "profiles/dpx_peaks.json (ba_microbench on B200)"}
            if traffic is not None:  # the same ncu capture: how busy the integer/DPX pipe really is
                roof["dpx"]["ncu_alu_pipe_pct"] = ent.get("alu_pipe_pct")
                roof["dpx"]["ncu_issue_active_pct"] = ent.get("issue_active_pct")
                roof["dpx"]["ncu_thread_inst_per_cell"] = ent.get("inst_per_cell")
        line = {
            "metric": METRIC, "value": value, "unit": "GCUPS", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int16" if agg["pairs_fast16"] else "int32",
            "data": "synthetic",
            "config": {"workload": w.name, "pairs_per_gpu_per_step": w.n, "cells_per_gpu_per_step": cells_step,
                       "parallelism": f"independent pairs sharded over {world} GPU(s), no collective",
                       "l2": "inputs + traceback codes per step exceed the 126 MB L2",
                       "seed": "splitmix64 0xB10600+config (+rank<<20)",
                       "host_plan": "work lists cached across same-shaped batches (lengths only, never results)"},
            "roofline": roof, "clocks": clocks,
            "breakdown_ms_per_step": {"fill": agg["ms_fill"] / args.steps, "trace": agg["ms_trace"] / args.steps,
                                      "d2h": agg["ms_d2h"] / args.steps, "host_plan_scatter": agg["ms_host"] / args.steps,
                                      "pairs_on_dpx16_path": agg["pairs_fast16"] // args.steps},
            "gpu_launches": int(agg["gpu_launches"]),
        }
        if e2e:
            line["e2e"] = e2e
        if not args.no_cpu_baseline and world == 1:
            threads = os.cpu_count() or 1
            sample = args.cpu_sample or max(64, 2000 * threads)
            g, n, c, dt = cpu_port_gcups(w, sample, threads)
            line["cpu_baseline"] = {"value": g, "unit": "GCUPS", "cores": threads, "kind": "port",
                                    "sample": f"first {n} pairs of the step ({c} cells), C port of the Go path "
                                              f"(full int64 table per pair), {threads} pthreads, {dt:.1f} s"}
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
