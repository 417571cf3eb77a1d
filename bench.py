#!/usr/bin/env python
"""bench.py -- STFT + 14-feature frames/s (fft 2048, hop 512, Hann) on N B200s, beside the CPU port.

Contract (see the task statement): one JSON line on stdout from rank 0.
  value      whole-job frames/s with inputs resident in HBM (device-timed, CUDA events, max over ranks)
  e2e        same metric through the host-pointer C-ABI call (pinned host buffers; H2D + kernel + D2H timed), with the
             platform's own pinned H2D rate measured beside it (e2e.platform_h2d_gbs) and the pageable-memory numbers a
             Rust `Vec<f64>` caller would see (e2e_pageable)
  roofline   HBM roofline of the fused K1 kernel: algorithmic bytes / measured launch duration / measured peak
  cpu_baseline  the C port of the reference's threaded analyzer on this box's host cores (bounded sample)
`--impl reference` times that CPU port alone (the reference is Rust and cannot be built in this image); it loads
oracle/libaus_oracle.so only -- never the CUDA library.

Workload at any N ("config 5 shard"): every rank holds FILES_PER_GPU synthetic 1-hour mono 44.1 kHz files
(0.5*sin + 0.1*uniform noise, counter-based seed 0, per-file frequency), generated on the device, f64, and analyses
all of them in one fused launch per step: fft_size 2048, hop 512, Hann, full 14-feature set.  Weak scaling: files are
independent, there is no data-path collective.

Other legs (each prints ONE JSON line of its own; none of them is the driver's default):
  --config 1..4   the other BASELINE.json configurations, device-timed with the same clock / L2 rules
  --corpus        sustained config-5 run on one GPU: generate -> analyse -> stream the feature rows out, three streams
  --multi         the product's single-process multi-GPU entry points (aus_stft_features_multi, aus_irstft_multi) over
                  the visible GPUs, strong scaling on config 4 (one 1-h file by frame range) and config 3 (1000 clips)
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SR = 44100
FFT, HOP, WINDOW = 2048, 512, 2          # Hann
FILE_SAMPLES = 3600 * SR                 # 1 h
METRIC = "stft_feature_frames_per_sec_fft2048_hop512"


def _frames(n, N=FFT, hop=HOP):
    return 1 if n <= N else -(-(n - N) // hop) + 1


def file_freq(file_id: int) -> float:
    """Per-file sine frequency of the synthetic corpus: 110 Hz * 2^(k/12), k = file_id mod 61."""
    return 110.0 * 2.0 ** ((file_id % 61) / 12.0)


def workload_config(files: int) -> dict:
    """The `config` object both arms print (the reference arm times a bounded sample of the same workload)."""
    return {"workload": f"config-5 shard: {files} x 1-hour 44.1 kHz mono files per GPU, fft 2048 hop 512 Hann, "
                        f"14 features, one fused launch per step",
            "files_per_gpu": files, "frames_per_step_per_gpu": _frames(FILE_SAMPLES) * files, "l2": "inputs_larger_than_l2",
            "sharding": "by file, no collective"}


def hbm_peak():
    pp = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pp):
        return float(json.load(open(pp))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# --------------------------------------------------------------------------------------------- clocks sampler
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        busy = [s for s in sm if s > 0.5 * max(sm)] or sm
        return {"sm_mhz": statistics.median(busy), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------------- CPU port
def _oracle_lib():
    """oracle/libaus_oracle.so, compiled here with gcc when it is missing or stale.  Nothing of the product is loaded."""
    import ctypes
    src = os.path.join(ROOT, "oracle", "aus_oracle.c")
    so = os.path.join(ROOT, "oracle", "libaus_oracle.so")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.run(["gcc", "-O3", "-fPIC", "-std=c11", "-fno-fast-math", "-shared", "-o", so, src, "-lm", "-lpthread"],
                       check=True, cwd=os.path.join(ROOT, "oracle"))
    return ctypes.CDLL(so), ctypes


def cpu_port_rate(seconds_audio: float, threads: int, mode: int, repeats: int = 1):
    """frames/s of oracle/aus_oracle.c (`mp.rs`-shaped: mode 0 = serial STFT + atan2 polar, analyzer on the pool;
    mode 1 = every stage parallel, no atan2) on `seconds_audio` of the synthetic signal."""
    import numpy as np
    lib, ct = _oracle_lib()
    n = int(seconds_audio * SR)
    x = np.empty(n)
    pd = ct.POINTER(ct.c_double)
    lib.aus_oracle_synth(x.ctypes.data_as(pd), ct.c_size_t(n), ct.c_size_t(0), ct.c_uint32(SR), ct.c_double(440.0),
                         ct.c_uint64(0), ct.c_uint64(0))
    F = _frames(n)
    out = np.empty((F, 14))
    best = float("inf")
    for _ in range(repeats):
        t0 = time.perf_counter()
        rc = lib.aus_oracle_stft_features_mt(x.ctypes.data_as(pd), ct.c_size_t(n), ct.c_size_t(FFT), ct.c_size_t(HOP),
                                             ct.c_int(WINDOW), ct.c_uint32(SR), ct.c_size_t(threads), ct.c_int(mode),
                                             out.ctypes.data_as(pd))
        best = min(best, time.perf_counter() - t0)
        assert rc == 0
    return F / best, F, best


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    # one step = one file of the workload when that keeps the run within a few minutes (a 1-hour file is ~15 s of CPU
    # work), otherwise a bounded prefix of it; the port's rate does not depend on the length (measured: 300 s vs 3600 s
    # differ by < 7 %)
    secs = 3600.0 if args.steps <= 8 else max(300.0, min(3600.0, 30000.0 / args.steps))
    for _ in range(args.warmup):
        cpu_port_rate(10.0, cores, 0)
    t0 = time.perf_counter()
    frames = 0
    for _ in range(args.steps):
        r, F, dt = cpu_port_rate(secs, cores, 0)
        frames += F
    dt = time.perf_counter() - t0
    value = frames / dt
    what = "one whole 1-hour file" if secs == 3600.0 else f"the first {secs:.0f} s of a 1-hour file"
    sample = (f"{what} per step ({_frames(int(secs * SR))} frames), synthetic mono 44.1 kHz, fft 2048 hop 512 Hann, reference structure "
              f"(serial rstft + atan2 polar, analyzer on {cores} threads), radix-4 FFT port")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.files),
        "audio_seconds_per_sec": value * HOP / SR,
        "cpu_baseline": {"value": value, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# --------------------------------------------------------------------------------------------- GPU plumbing
class Rig:
    """Process-wide GPU plumbing of the GPU arm: rank / device, NCCL for the barrier and the MAX reduction, context."""

    def __init__(self):
        import torch
        import torch.distributed as dist
        import __graft_entry__ as g
        g.build()
        import aus_rust_b200 as aus
        self.torch, self.dist, self.aus = torch, dist, aus
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        self.affinity = None
        try:        # run (and allocate the pinned host buffers) on the CPUs next to this rank's GPU: e2e is a PCIe / host-memory number
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(torch.cuda._get_nvml_device_index(self.local) if hasattr(torch.cuda, "_get_nvml_device_index") else self.local)
            pynvml.nvmlDeviceSetCpuAffinity(h)
            self.affinity = len(os.sched_getaffinity(0))
        except Exception:
            self.affinity = None
        try:
            self.numa_nodes = len([d for d in os.listdir("/sys/devices/system/node") if d.startswith("node")])
        except Exception:
            self.numa_nodes = None
        if self.world > 1:
            # NCCL prints its version banner on stdout when the first communicator comes up; the contract is ONE JSON line
            # on stdout, so stdout points at stderr until the communicator exists
            sys.stdout.flush()
            saved = os.dup(1)
            os.dup2(2, 1)
            try:
                dist.init_process_group("nccl", device_id=self.dev)
                dist.barrier()
                torch.cuda.synchronize()
            finally:
                sys.stdout.flush()
                os.dup2(saved, 1)
                os.close(saved)
        self.ctx = aus.Context(self.local)
        self.ctx.use_torch_stream()

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x: float) -> float:
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def timed_steps(self, step, steps: int, warmup: int, flush=None):
        """W warm-up steps, then exactly `steps` steps bracketed by barrier + synchronize; CUDA events on the launching
        stream around every step; returns (max-over-ranks total ms, per-step ms list, launches, clocks or None).
        `flush`: a callable run before every timed step outside its event pair (L2 flush for inputs smaller than L2)."""
        torch = self.torch
        for _ in range(max(warmup, 3)):
            step()
        self.barrier()
        sampler = ClockSampler(self.local)
        if self.rank == 0:
            sampler.start()
        l0 = self.ctx.launch_count
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        for i in range(steps):
            if flush is not None:
                flush()
            ev[i][0].record()
            step()
            ev[i][1].record()
        self.barrier()
        launches = self.ctx.launch_count - l0
        clocks = sampler.stop() if self.rank == 0 else None
        per = [a.elapsed_time(b) for a, b in ev]
        total = ev[0][0].elapsed_time(ev[-1][1]) if flush is None else sum(per)
        return self.max_over_ranks(total), per, launches, clocks

    def close(self):
        if self.world > 1:
            self.dist.destroy_process_group()


def h2d_probe(rig: Rig, nbytes: int, reps: int = 3):
    """What the platform gives a plain pinned cudaMemcpyAsync of the same size, all ranks copying at once: the ceiling
    of any end-to-end number that has to move `nbytes` per step to the device.  GB/s per rank (slowest rank)."""
    torch = rig.torch
    h = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    d = torch.empty(nbytes, dtype=torch.uint8, device=rig.dev)
    d.copy_(h, non_blocking=True)
    rig.barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    dt = rig.max_over_ranks(time.perf_counter() - t0)
    del h, d
    return nbytes * reps / dt / 1e9


# --------------------------------------------------------------------------------------------- GPU arm, default workload
def run_ours(args):
    import numpy as np
    rig = Rig()
    torch, ctx, aus, dev, rank, world = rig.torch, rig.ctx, rig.aus, rig.dev, rig.rank, rig.world
    prec = args.precision
    tdt = torch.float64 if prec == "f64" else torch.float32
    esz = 8 if prec == "f64" else 4
    if args.direct_load is not None:
        ctx.set_direct_load(bool(args.direct_load))

    files = args.files
    lens = np.full(files, FILE_SAMPLES, dtype=np.uint64)
    offs = np.arange(files, dtype=np.uint64) * np.uint64(FILE_SAMPLES)
    frames_per_file = _frames(FILE_SAMPLES)
    frames = frames_per_file * files
    d_audio = torch.empty(files * FILE_SAMPLES, dtype=tdt, device=dev)
    for i in range(files):
        fid = rank * files + i
        ctx.synth_dev(d_audio[i * FILE_SAMPLES:(i + 1) * FILE_SAMPLES], 0, SR, file_freq(fid), 0, fid, prec)
    d_feat = torch.empty((frames, 14), dtype=tdt, device=dev)
    torch.cuda.synchronize()

    def step():
        ctx.stft_features_batch_dev(d_audio, offs, lens, FFT, HOP, WINDOW, SR, prec, d_feat)

    max_ms, kern_ms, launches, clocks = rig.timed_steps(step, args.steps, args.warmup)
    value = frames * world * args.steps / (max_ms * 1e-3)

    # ---- property check on the full-size result (size-independent): finite, roll-offs ordered and on the bin grid
    feat = d_feat[:: max(1, frames // 4096)].double().cpu().numpy()
    assert np.isfinite(feat).all(), "non-finite features in the benchmark output"
    assert (np.diff(feat[:, 6:10], axis=1) >= 0).all(), "roll-off points must be non-decreasing"
    bins = feat[:, 6:10] / (SR / FFT)
    assert np.abs(bins - np.round(bins)).max() < 1e-3, "roll-off points must sit on the rfftfreq grid"

    # ---- the other precision of the same workload (extra keys; BASELINE config 5 asks for f64 and f32)
    other = None
    if not args.no_other_precision:
        op = "f32" if prec == "f64" else "f64"
        odt = torch.float32 if op == "f32" else torch.float64
        osz = 4 if op == "f32" else 8
        ofiles = min(files, 2)
        oa = d_audio[: ofiles * FILE_SAMPLES].to(odt)
        of = torch.empty((frames_per_file * ofiles, 14), dtype=odt, device=dev)

        def ostep():
            ctx.stft_features_batch_dev(oa, offs[:ofiles], lens[:ofiles], FFT, HOP, WINDOW, SR, op, of)

        oms, _, _, _ = rig.timed_steps(ostep, max(3, min(args.steps, 10)), 3)
        osteps = max(3, min(args.steps, 10))
        ov = frames_per_file * ofiles * world * osteps / (oms * 1e-3)
        obytes = ofiles * FILE_SAMPLES * osz + frames_per_file * ofiles * 14 * osz
        other = {"dtype": op, "value": ov, "unit": "frames/s", "files_per_gpu": ofiles, "steps": osteps,
                 "hbm_frac": obytes * osteps * world / (oms * 1e-3) / 1e9 / (hbm_peak()[0] * world)}
        del oa, of

    # ---- e2e: the host-pointer C-ABI call on pinned host buffers (H2D + K1 + D2H inside the timed region)
    e2e_files = min(files, args.e2e_files)
    h_audio = torch.empty(e2e_files * FILE_SAMPLES, dtype=torch.float64, pin_memory=True)
    h_audio.copy_(d_audio[: e2e_files * FILE_SAMPLES].double())
    e2e_frames = frames_per_file * e2e_files
    h_out = torch.empty((e2e_frames, 14), dtype=torch.float64, pin_memory=True)
    a_np, o_np = h_audio.numpy(), h_out.numpy()
    e2e_steps = max(2, min(args.steps, 5))

    def e2e_time(call):
        for _ in range(2):
            call()
        rig.barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            call()
        torch.cuda.synchronize()
        return e2e_frames * world * e2e_steps / rig.max_over_ranks(time.perf_counter() - t0)

    e2e_value = e2e_time(lambda: ctx.stft_features_batch(a_np, offs[:e2e_files], lens[:e2e_files], FFT, HOP, WINDOW, SR, prec, out=o_np))
    chk = d_feat[:1000].double().cpu().numpy()       # the e2e result must be the numbers of the device-resident run
    assert np.allclose(o_np[:1000], chk, rtol=1e-6 if prec == "f32" else 1e-12, atol=0, equal_nan=True)
    h2d_gbs = h2d_probe(rig, e2e_files * FILE_SAMPLES * 8)
    e2e_bytes = e2e_files * FILE_SAMPLES * 8 + e2e_frames * 14 * 8
    e2e_gbs = e2e_bytes * e2e_value / (e2e_frames * world) / 1e9          # per rank, bytes moved per second by the call
    # ---- the same call on PAGEABLE host memory (what a Rust Vec<f64> is), with and without page-locking it for the call
    pageable = None
    try:
        p_np = np.empty(e2e_files * FILE_SAMPLES, dtype=np.float64)
        p_np[:] = a_np
        po_np = np.empty((e2e_frames, 14), dtype=np.float64)
        res = {}
        for mode, name in ((0, "driver_staged"), (1, "host_register"), (2, "pinned_ring_parallel_memcpy")):
            ctx.set_host_register(mode)
            res[name] = e2e_time(lambda: ctx.stft_features_batch(p_np, offs[:e2e_files], lens[:e2e_files], FFT, HOP, WINDOW, SR, prec, out=po_np))
        ctx.set_host_register(-1)
        assert np.array_equal(po_np[:1000], o_np[:1000], equal_nan=True)
        ctx.set_host_register(-1)
        auto = e2e_time(lambda: ctx.stft_features_batch(p_np, offs[:e2e_files], lens[:e2e_files], FFT, HOP, WINDOW, SR, prec, out=po_np))
        pageable = {"value": auto, "unit": "frames/s", **res, "host_dtype": "f64 (pageable numpy array = a Rust Vec<f64>)",
                    "note": "value = the default policy (aus_ctx_set_host_register -1); the three modes are measured beside it"}
        del p_np, po_np
    except Exception as exc:                       # extra information only: never fail the bench line over it
        pageable = {"error": str(exc)[:200]}
    # ---- the same call fed with decoded 16-bit PCM instead of f64 samples (`sample as f64 / INTMAX` of
    #      audiofile.rs:296-329 done on the device): 2 bytes per sample over PCIe instead of 8.  Extra information, not `e2e`.
    e2e_pcm = None
    try:
        h_pcm = torch.empty(e2e_files * FILE_SAMPLES, dtype=torch.int16, pin_memory=True)
        h_pcm.copy_((d_audio[: e2e_files * FILE_SAMPLES].double() * 32767.0).round().clamp(-32768, 32767).to(torch.int16))
        p16 = h_pcm.numpy()
        v = e2e_time(lambda: ctx.stft_features_batch_pcm(p16, offs[:e2e_files], lens[:e2e_files], FFT, HOP, WINDOW, SR, prec, out=o_np))
        e2e_pcm = {"value": v, "unit": "frames/s", "h2d_bytes_per_step": e2e_files * FILE_SAMPLES * 2,
                   "d2h_bytes_per_step": e2e_frames * 14 * 8,
                   "host_dtype": "int16 PCM (pinned), widened on the device like audiofile.rs:296-329"}
    except Exception as exc:
        e2e_pcm = {"error": str(exc)[:200]}

    if rank == 0:
        peak, peak_src = hbm_peak()
        alg_bytes = files * FILE_SAMPLES * esz + frames * 14 * esz          # every sample once, every feature once
        k_ms = statistics.mean(kern_ms)
        achieved = alg_bytes / (k_ms * 1e-3) / 1e9
        traffic, tj = None, {}
        tp = os.path.join(ROOT, "profiles", "k1_traffic.json")
        if os.path.exists(tp):
            try:
                tj = json.load(open(tp))
                traffic = tj.get(f"dram_bytes_per_frame_{prec}") and tj[f"dram_bytes_per_frame_{prec}"] * frames
            except Exception:
                traffic = None
        # what actually bounds the fused f64 kernel (extra key; the contract's `roofline` stays the HBM one): FP64 issue.
        # 64 FP64 lanes per SM -> one FP64 warp-instruction per 2 cycles per scheduler; the per-frame count is ncu's.
        compute = None
        if prec == "f64" and isinstance(tj, dict) and tj.get("fp64_warp_inst_per_frame_f64") and isinstance(clocks, dict):
            sm_hz = 1e6 * float(clocks.get("sm_mhz") or clocks.get("sm_max_mhz") or 1965.0)
            peak_wi = torch.cuda.get_device_properties(dev).multi_processor_count * 4 * sm_hz / 2.0
            ach_wi = tj["fp64_warp_inst_per_frame_f64"] * frames / (k_ms * 1e-3)
            compute = {"bound": "fp64 pipe", "achieved": ach_wi / 1e12, "peak": peak_wi / 1e12, "unit": "T warp-instr/s",
                       "frac": ach_wi / peak_wi, "fp64_warp_inst_per_frame": tj["fp64_warp_inst_per_frame_f64"],
                       "peak_source": "SMs x 4 schedulers x SM clock under load / 2 (16 FP64 lanes per scheduler)"}
        cores = os.cpu_count() or 1
        cpu = None
        if world == 1 and not args.no_cpu:
            r0, F0, dt0 = cpu_port_rate(3600.0, cores, 0)                   # one of the workload's 1-hour files: ~15 s
            r1, F1, dt1 = cpu_port_rate(3600.0, cores, 1)
            cpu = {"value": r0, "unit": "frames/s", "cores": cores, "kind": "port",
                   "sample": f"one 1-hour synthetic mono 44.1 kHz file ({F0} frames), fft 2048 hop 512 Hann, C port of mp.rs structure "
                             f"(serial rstft + atan2 polar, analyzer on {cores} threads, radix-4 FFT), {dt0:.1f} s",
                   "all_stages_parallel_value": r1}
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": max_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": prec, "data": "synthetic",
            "config": workload_config(files),
            "audio_seconds_per_sec": value * HOP / SR,
            "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": e2e_files * FILE_SAMPLES * 8,
                    "d2h_bytes_per_step": e2e_frames * 14 * 8, "files_per_step": e2e_files, "steps": e2e_steps,
                    "host_dtype": "f64 (pinned)", "rank_cpu_affinity": rig.affinity, "numa_nodes": rig.numa_nodes,
                    "platform_h2d_gbs": h2d_gbs, "call_gbs_per_rank": e2e_gbs, "frac_of_platform_h2d": e2e_gbs / h2d_gbs,
                    "note": "platform_h2d_gbs = a bare pinned cudaMemcpyAsync of the same input, all ranks at once (slowest rank): "
                            "the ceiling of any call that must move 8 bytes per sample to the device"},
            "e2e_pageable": pageable,
            "e2e_pcm16": e2e_pcm,
            "other_precision": other,
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src, "kernel": "stft_kernel<double, 11, 0, 1>" if prec == "f64" else "stft_kernel<float, 11, 1, 1>",
                         "kernel_ms": k_ms, "algorithmic_bytes_per_launch": alg_bytes,
                         "note": "fused f64 path is bound by per-warp dependency chains at the occupancy the register file allows (FP64 pipe, "
                                 "issue slots and L1/shared pipe each ~50-65 % busy), not by HBM (DESIGN.md section 5, "
                                 "profiles/r2_k1_pc_experiment.md); the spectra-writing kernel of the same path runs at ~73 %"},
            "roofline_fp64": compute,
            "cpu_baseline": cpu,
            "clocks": clocks,
        }))
    rig.close()


# --------------------------------------------------------------------------------------------- GPU arm, BASELINE configs 1-4
def run_config(args):
    """One of the other BASELINE.json configurations, device-resident, same timing rules (W >= 3 warm-ups, CUDA events on
    the launching stream, L2 flushed between steps when the input is smaller than L2, clocks sampled)."""
    import numpy as np
    rig = Rig()
    torch, ctx, dev, world, rank = rig.torch, rig.ctx, rig.dev, rig.world, rig.rank
    prec = args.precision
    tdt = torch.float64 if prec == "f64" else torch.float32
    esz = 8 if prec == "f64" else 4
    peak, peak_src = hbm_peak()
    cfg = args.config
    flushbuf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # 256 MB > the 126 MB L2

    def flush():
        flushbuf.add_(1)

    def synth(n_files, n):
        d = torch.empty(n_files * n, dtype=tdt, device=dev)
        for i in range(n_files):
            fid = rank * n_files + i
            ctx.synth_dev(d[i * n:(i + 1) * n], 0, SR, file_freq(fid), 0, fid, prec)
        return d

    extra = {}
    if cfg in (1, 3, 4):
        N, hop, n_files, n = {1: (2048, 512, 1, 60 * SR), 3: (1024, 256, 1000, 30 * SR), 4: (16384, 4096, 1, 3600 * SR)}[cfg]
        d = synth(n_files, n)
        F = _frames(n, N, hop) * n_files
        lens = np.full(n_files, n, dtype=np.uint64)
        offs = np.arange(n_files, dtype=np.uint64) * np.uint64(n)
        feat = torch.empty((F, 14), dtype=tdt, device=dev)

        def step():
            ctx.stft_features_batch_dev(d, offs, lens, N, hop, WINDOW, SR, prec, feat)

        alg = n_files * n * esz + F * 14 * esz
        small = n_files * n * esz < (200 << 20)
        metric = f"stft_feature_frames_per_sec_fft{N}_hop{hop}"
        workload = {1: "config 1: 60 s mono 44.1 kHz, fft 2048 hop 512 Hann, 14 features, one call",
                    3: "config 3: 1000 x 30 s clips, fft 1024 hop 256 Hann, features only, one launch",
                    4: "config 4: 1 h mono, fft 16384 hop 4096 Hann, 14 features incl. sub-band slopes and roll-offs"}[cfg]
    else:
        # config 2: rstft -> irstft of 10 min stereo, fft 4096 hop 1024; one step = both channels forward + one batched inverse
        N, hop, n = 4096, 1024, 600 * SR
        d = synth(2, n)
        Fc, K = _frames(n, N, hop), N // 2 + 1
        F = 2 * Fc
        lens = np.full(2, n, dtype=np.uint64)
        offs = np.arange(2, dtype=np.uint64) * np.uint64(n)
        spec = torch.empty((F, K, 2), dtype=tdt, device=dev)
        L = N + hop * (Fc - 1)
        y = torch.empty(2 * L, dtype=tdt, device=dev)
        nf = np.array([Fc, Fc], dtype=np.uint64)

        def step():
            ctx.stft_features_batch_dev(d, offs, lens, N, hop, 3, SR, prec, None, spec)
            ctx.irstft_batch_dev(spec, nf, N, hop, 3, prec, y, True)

        alg = 2 * n * esz + 2 * F * K * 2 * esz + 3 * 2 * L * esz          # in + spectra written and read + out written, rescaled
        small = False
        metric = "rstft_irstft_roundtrip_frames_per_sec_fft4096_hop1024"
        workload = "config 2: rstft -> irstft of 10 min stereo 44.1 kHz, fft 4096 hop 1024 Hamming, both channels per step"
    steps = args.steps
    max_ms, per, launches, clocks = rig.timed_steps(step, steps, args.warmup, flush if small else None)
    value = F * world * steps / (max_ms * 1e-3)
    if cfg == 2:
        x0 = d[:n].double()
        body = slice(N, n - 2 * N)
        gain = float((y[:L][body].double() * x0[body]).sum() / (x0[body] * x0[body]).sum())
        extra["reconstruction_max_abs_error_body"] = float((y[:L][body].double() / gain - x0[body]).abs().max())
        extra["peak_after_normalisation"] = float(y[:L].max())
    if rank == 0:
        k_ms = statistics.mean(per)
        ach = alg / (k_ms * 1e-3) / 1e9
        print(json.dumps({
            "metric": metric, "value": value, "unit": "frames/s", "n_gpus": world, "steps": steps, "warmup": max(args.warmup, 3),
            "ms_per_step": max_ms / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": prec,
            "data": "synthetic",
            "config": {"workload": workload, "l2": "flushed_between_steps (256 MB write)" if small else "inputs_larger_than_l2",
                       "baseline_config": cfg},
            "audio_seconds_per_sec": value * hop / SR,
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None,
                         "peak_source": peak_src, "kernel_ms": k_ms, "algorithmic_bytes_per_launch": alg},
            "clocks": clocks, **extra,
        }))
    rig.close()


# --------------------------------------------------------------------------------------------- single-process multi-GPU leg
def run_multi(args):
    """The product's own multi-GPU entry points from ONE process over the visible GPUs (host threads, shard plan, ordered
    gather; no torchrun, no collective): strong scaling on a fixed problem at 1, 2, 4, ... devices.  The calls take HOST
    buffers (pageable, page-locked by the call), so the numbers include H2D and D2H."""
    import numpy as np
    import torch
    import __graft_entry__ as g
    g.build()
    import aus_rust_b200 as aus
    have = torch.cuda.device_count()
    counts = [c for c in (1, 2, 4, 8) if c <= min(have, args.gpus if args.gpus > 1 else have)]
    ctx = aus.Context(0)

    def host_signal(n_files, n, pinned):
        """synthetic files as one host array: pageable (a numpy array = what a Rust Vec is) or pinned"""
        d = torch.empty(n_files * n, dtype=torch.float64, device="cuda:0")
        for i in range(n_files):
            ctx.synth_dev(d[i * n:(i + 1) * n], 0, SR, file_freq(i), 0, i, "f64")
        torch.cuda.synchronize()
        if pinned:
            h = torch.empty(n_files * n, dtype=torch.float64, pin_memory=True)
            h.copy_(d)
            return h.numpy(), h
        return d.cpu().numpy(), None

    def best_of(fn, reps=3):
        fn()
        best = float("inf")
        for _ in range(reps):
            t0 = time.perf_counter()
            fn()
            best = min(best, time.perf_counter() - t0)
        return best

    out = {"metric": "multi_gpu_strong_scaling_frames_per_sec", "unit": "frames/s", "devices_visible": have, "n_gpus": max(counts),
           "entry_points": "aus_stft_features_multi / aus_irstft_multi (one process, one host thread + context per device)",
           "data": "synthetic", "higher_is_better": True, "workloads": {}}
    for pinned in (False, True):
        tag = "pinned_host" if pinned else "pageable_host"
        # config 4: one 1-hour file, fft 16384 / hop 4096, split by frame range with an fft_size - hop sample halo
        x, keep = host_signal(1, 3600 * SR, pinned)
        F4 = _frames(x.size, 16384, 4096)
        ref, res = None, {}
        for c in counts:
            devs = list(range(c))
            t = best_of(lambda: aus.stft_features_multi(devs, x, [0], [x.size], 16384, 4096, WINDOW, SR))
            got = aus.stft_features_multi(devs, x, [0], [x.size], 16384, 4096, WINDOW, SR)
            if ref is None:
                ref = got
            assert np.array_equal(got, ref, equal_nan=True), "sharded result differs from the single-device one"
            res[str(c)] = F4 / t
        out["workloads"].setdefault("config4_one_1h_file_fft16384_hop4096_by_frame_range", {"frames": F4, "h2d_bytes": x.size * 8,
                                    "bit_identical_across_device_counts": True})[tag] = res
        del x, keep
        # config 3: 1000 x 30 s clips, fft 1024 / hop 256, whole files per device
        n = 30 * SR
        x, keep = host_signal(1000, n, pinned)
        lens = np.full(1000, n, dtype=np.uint64)
        offs = np.arange(1000, dtype=np.uint64) * np.uint64(n)
        F3 = _frames(n, 1024, 256) * 1000
        ref, res = None, {}
        for c in counts:
            devs = list(range(c))
            t = best_of(lambda: aus.stft_features_multi(devs, x, offs, lens, 1024, 256, WINDOW, SR))
            got = aus.stft_features_multi(devs, x, offs, lens, 1024, 256, WINDOW, SR)
            if ref is None:
                ref = got
            assert np.array_equal(got, ref, equal_nan=True)
            res[str(c)] = F3 / t
        out["workloads"].setdefault("config3_1000x30s_fft1024_hop256_by_file", {"frames": F3, "h2d_bytes": x.size * 8,
                                    "bit_identical_across_device_counts": True})[tag] = res
        del x, keep
    # config 2: irstft of a 10-minute stereo pair, fft 4096 / hop 1024, by frame range (by channel at 2 devices); pageable
    n = 600 * SR
    xs, _ = host_signal(2, n, False)
    S = [ctx.rstft(xs[i * n:(i + 1) * n], 4096, 1024, 3) for i in range(2)]
    Fc = S[0].shape[0]
    ref, res = None, {}
    for c in counts:
        devs = list(range(c))
        t = best_of(lambda: aus.irstft_multi(devs, S, 4096, 1024, 3), reps=2)
        got = aus.irstft_multi(devs, S, 4096, 1024, 3)
        if ref is None:
            ref = got
        err = max(float(np.max(np.abs(a - b))) for a, b in zip(got, ref))
        assert err <= 1e-12, err                     # only the peak factor's last bits may differ between shardings
        res[str(c)] = 2 * Fc / t
    out["workloads"]["config2_irstft_10min_stereo_fft4096_hop1024"] = {"frames": 2 * Fc, "pageable_host": res,
                                                                       "h2d_bytes": int(sum(s.nbytes for s in S)),
                                                                       "note": "includes the Python-side concatenation of the two spectrograms"}
    aus.multi_shutdown()
    print(json.dumps(out))


# --------------------------------------------------------------------------------------------- sustained corpus leg
def run_corpus(args):
    """SURVEY 7 hard part 5 / BASELINE config 5 as a sustained run on one GPU: a stream of 1-hour files is generated on the
    device, analysed (one fused launch per file) and its feature rows are streamed to pinned host memory, with the three
    stages on three streams and double buffers (generation of file i+1 and the D2H of file i-1 overlap the analysis of
    file i).  Reported with and without the generation stage; the corpus rate of a node is this times its GPUs."""
    import numpy as np
    import torch
    import __graft_entry__ as g
    g.build()
    import aus_rust_b200 as aus
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    prec = args.precision
    tdt = torch.float64 if prec == "f64" else torch.float32
    n_files = args.corpus_files
    F = _frames(FILE_SAMPLES)
    s_gen, s_cmp, s_out = (torch.cuda.Stream(device=dev) for _ in range(3))
    ctx_gen, ctx_cmp = aus.Context(0), aus.Context(0)
    ctx_gen.set_stream(s_gen.cuda_stream)
    ctx_cmp.set_stream(s_cmp.cuda_stream)
    audio = [torch.empty(FILE_SAMPLES, dtype=tdt, device=dev) for _ in range(2)]
    feat = [torch.empty((F, 14), dtype=tdt, device=dev) for _ in range(2)]
    host = [torch.empty((F, 14), dtype=tdt, pin_memory=True) for _ in range(2)]
    lens, offs = np.array([FILE_SAMPLES], dtype=np.uint64), np.array([0], dtype=np.uint64)
    checks = []

    def run(generate: bool):
        ev_gen = [torch.cuda.Event() for _ in range(2)]
        ev_cmp = [torch.cuda.Event() for _ in range(2)]
        ev_out = [torch.cuda.Event() for _ in range(2)]
        if not generate:                                   # analysis + output stream only: both buffers hold a file already
            for b in range(2):
                ctx_gen.synth_dev(audio[b], 0, SR, file_freq(b), 0, b, prec)
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(n_files + 1):
            b = i & 1
            if i < n_files and generate:
                if i >= 2:
                    s_gen.wait_event(ev_cmp[b])            # the analysis of file i-2 has released audio[b]
                ctx_gen.synth_dev(audio[b], 0, SR, file_freq(i), 0, i, prec)
                ev_gen[b].record(s_gen)
            if i >= 1:                                     # analyse file i-1 while file i is being generated
                j, bj = i - 1, (i - 1) & 1
                if generate:
                    s_cmp.wait_event(ev_gen[bj])
                if j >= 2:
                    s_cmp.wait_event(ev_out[bj])           # the D2H of file j-2 has released feat[bj]
                ctx_cmp.stft_features_batch_dev(audio[bj], offs, lens, FFT, HOP, WINDOW, SR, prec, feat[bj])
                ev_cmp[bj].record(s_cmp)
                s_out.wait_event(ev_cmp[bj])
                with torch.cuda.stream(s_out):
                    host[bj].copy_(feat[bj], non_blocking=True)
                ev_out[bj].record(s_out)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        checks.append(float(host[(n_files - 1) & 1].double().abs().sum()))
        return n_files * F / dt

    run(True)                                              # warm-up (tables, allocations)
    with_gen = run(True)
    without = run(False)
    esz = 8 if prec == "f64" else 4
    print(json.dumps({
        "metric": "corpus_sustained_frames_per_sec_fft2048_hop512", "unit": "frames/s", "n_gpus": 1, "dtype": prec, "data": "synthetic",
        "higher_is_better": True, "files": n_files, "hours_of_audio": n_files,
        "generate_analyse_stream": with_gen, "analyse_stream": without,
        "audio_hours_per_hour_per_gpu": {"generate_analyse_stream": with_gen * HOP / SR, "analyse_stream": without * HOP / SR},
        "feature_d2h_gbs": without * 14 * esz / 1e9,
        "ten_thousand_hour_corpus_seconds_on_8_gpus": {"generate_analyse_stream": 10000 * 3600 / (8 * with_gen * HOP / SR),
                                                       "analyse_stream": 10000 * 3600 / (8 * without * HOP / SR)},
        "config": {"workload": "config 5 sustained: 1-hour files through double buffers, three streams (generate / fused K1 + finalize / "
                               "feature rows to pinned host memory)"},
        "finite": bool(np.isfinite(checks).all()),
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="f64", choices=["f64", "f32"])
    ap.add_argument("--files", type=int, default=4, help="1-hour files resident per GPU")
    ap.add_argument("--e2e-files", type=int, default=2)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-other-precision", action="store_true")
    ap.add_argument("--config", type=int, default=5, choices=[1, 2, 3, 4, 5], help="BASELINE.json configuration (5 = the metric's)")
    ap.add_argument("--multi", action="store_true", help="single-process multi-GPU strong-scaling leg")
    ap.add_argument("--corpus", action="store_true", help="sustained generate -> analyse -> stream-out leg on one GPU")
    ap.add_argument("--corpus-files", type=int, default=48, help="1-hour files of the --corpus leg")
    ap.add_argument("--direct-load", type=int, default=None, help="experiment: 1 = global->register input loads, 0 = TMA staging")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.multi:
        run_multi(args)
    elif args.corpus:
        run_corpus(args)
    elif args.config != 5:
        run_config(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
