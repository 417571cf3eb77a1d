#!/usr/bin/env python
"""bench.py -- STFT + 14-feature frames/s (fft 2048, hop 512, Hann) on N B200s, beside the CPU port.

Contract (see the task statement): one JSON line on stdout from rank 0.
  value      whole-job frames/s with inputs resident in HBM (device-timed, CUDA events, max over ranks)
  e2e        same metric through the host-pointer C-ABI call (pinned host buffers; H2D + kernel + D2H timed)
  roofline   HBM roofline of the fused K1 kernel: algorithmic bytes / measured launch duration / measured peak
  cpu_baseline  the C port of the reference's threaded analyzer on this box's host cores (bounded sample)
`--impl reference` times that CPU port alone (the reference is Rust and cannot be built in this image).

Workload at any N ("config 5 shard"): every rank holds FILES_PER_GPU synthetic 1-hour mono 44.1 kHz files
(0.5*sin + 0.1*uniform noise, counter-based seed 0, per-file frequency), generated on the device, f64, and analyses
all of them in one fused launch per step: fft_size 2048, hop 512, Hann, full 14-feature set.  Weak scaling: files are
independent, there is no data-path collective.
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
sys.path.insert(0, os.path.join(ROOT, "oracle"))

SR = 44100
FFT, HOP, WINDOW = 2048, 512, 2          # Hann
FILE_SAMPLES = 3600 * SR                 # 1 h
METRIC = "stft_feature_frames_per_sec_fft2048_hop512"


def _frames(n, N=FFT, hop=HOP):
    return 1 if n <= N else -(-(n - N) // hop) + 1


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
    import ctypes
    import __graft_entry__ as g
    g.build()
    lib = ctypes.CDLL(os.path.join(ROOT, "oracle", "libaus_oracle.so"))
    return lib, ctypes


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
    secs = 300.0                                  # bounded sample per step: 5 min of audio = 25 836 frames, ~1.5 s
    for _ in range(args.warmup):
        cpu_port_rate(10.0, cores, 0)
    t0 = time.perf_counter()
    frames = 0
    for _ in range(args.steps):
        r, F, dt = cpu_port_rate(secs, cores, 0)
        frames += F
    dt = time.perf_counter() - t0
    value = frames / dt
    sample = f"{secs:.0f} s synthetic mono 44.1 kHz per step, fft 2048 hop 512 Hann, reference structure (serial rstft + atan2 polar, analyzer on {cores} threads)"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "config-5 shard: 1-hour 44.1 kHz mono files, fft 2048 hop 512 Hann, 14 features (bounded CPU sample)"},
        "audio_seconds_per_sec": value * HOP / SR,
        "cpu_baseline": {"value": value, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# --------------------------------------------------------------------------------------------- GPU arm
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    import __graft_entry__ as g
    g.build()
    import aus_rust_b200 as aus
    import aus_oracle as orc

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = None
    try:        # run (and allocate the pinned host buffers) on the CPUs next to this rank's GPU: e2e is a PCIe / host-memory number
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(torch.cuda._get_nvml_device_index(local) if hasattr(torch.cuda, "_get_nvml_device_index") else local)
        pynvml.nvmlDeviceSetCpuAffinity(h)
        numa = len(os.sched_getaffinity(0))
    except Exception:
        numa = None
    if world > 1:
        # NCCL prints its version banner on stdout when the first communicator comes up; the contract is ONE JSON line
        # on stdout, so stdout points at stderr until the communicator exists
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    prec = args.precision
    tdt = torch.float64 if prec == "f64" else torch.float32
    esz = 8 if prec == "f64" else 4

    ctx = aus.Context(local)
    ctx.use_torch_stream()
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
        ctx.synth_dev(d_audio[i * FILE_SAMPLES:(i + 1) * FILE_SAMPLES], 0, SR, orc.file_freq(fid), 0, fid, prec)
    d_feat = torch.empty((frames, 14), dtype=tdt, device=dev)
    torch.cuda.synchronize()

    def step():
        ctx.stft_features_batch_dev(d_audio, offs, lens, FFT, HOP, WINDOW, SR, prec, d_feat)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = ctx.launch_count
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    ev[0].record()
    for i in range(args.steps):
        step()
        ev[i + 1].record()
    barrier()
    launches = ctx.launch_count - l0
    clocks = sampler.stop() if rank == 0 else None
    total_ms = ev[0].elapsed_time(ev[-1])
    kern_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]     # one K1 launch per step
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    max_ms = float(t.item())
    value = frames * world * args.steps / (max_ms * 1e-3)

    # ---- property check on the full-size result (size-independent): finite, roll-offs ordered and on the bin grid
    feat = d_feat[:: max(1, frames // 4096)].double().cpu().numpy()
    assert np.isfinite(feat).all(), "non-finite features in the benchmark output"
    assert (np.diff(feat[:, 6:10], axis=1) >= 0).all(), "roll-off points must be non-decreasing"
    bins = feat[:, 6:10] / (SR / FFT)
    assert np.abs(bins - np.round(bins)).max() < 1e-3, "roll-off points must sit on the rfftfreq grid"

    # ---- e2e: the host-pointer C-ABI call on pinned host buffers (H2D + K1 + D2H inside the timed region)
    e2e_files = min(files, args.e2e_files)
    h_audio = torch.empty(e2e_files * FILE_SAMPLES, dtype=torch.float64, pin_memory=True)
    h_audio.copy_(d_audio[: e2e_files * FILE_SAMPLES].double())
    e2e_frames = frames_per_file * e2e_files
    h_out = torch.empty((e2e_frames, 14), dtype=torch.float64, pin_memory=True)
    a_np, o_np = h_audio.numpy(), h_out.numpy()
    for _ in range(2):
        ctx.stft_features_batch(a_np, offs[:e2e_files], lens[:e2e_files], FFT, HOP, WINDOW, SR, prec, out=o_np)
    barrier()
    e2e_steps = max(2, min(args.steps, 5))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        ctx.stft_features_batch(a_np, offs[:e2e_files], lens[:e2e_files], FFT, HOP, WINDOW, SR, prec, out=o_np)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = e2e_frames * world * e2e_steps / float(te.item())
    # ---- the same call fed with decoded 16-bit PCM instead of f64 samples (`sample as f64 / INTMAX` of
    #      audiofile.rs:296-329 done on the device): 2 bytes per sample over PCIe instead of 8.  Extra information, not `e2e`.
    e2e_pcm = None
    try:
        h_pcm = torch.empty(e2e_files * FILE_SAMPLES, dtype=torch.int16, pin_memory=True)
        h_pcm.copy_((d_audio[: e2e_files * FILE_SAMPLES].double() * 32767.0).round().clamp(-32768, 32767).to(torch.int16))
        p_np = h_pcm.numpy()
        for _ in range(2):
            ctx.stft_features_batch_pcm(p_np, offs[:e2e_files], lens[:e2e_files], FFT, HOP, WINDOW, SR, prec, out=o_np)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            ctx.stft_features_batch_pcm(p_np, offs[:e2e_files], lens[:e2e_files], FFT, HOP, WINDOW, SR, prec, out=o_np)
        torch.cuda.synchronize()
        tp = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tp, op=dist.ReduceOp.MAX)
        e2e_pcm = {"value": e2e_frames * world * e2e_steps / float(tp.item()), "unit": "frames/s",
                   "h2d_bytes_per_step": e2e_files * FILE_SAMPLES * 2, "d2h_bytes_per_step": e2e_frames * 14 * 8,
                   "host_dtype": "int16 PCM (pinned), widened on the device like audiofile.rs:296-329"}
        ctx.stft_features_batch(a_np, offs[:e2e_files], lens[:e2e_files], FFT, HOP, WINDOW, SR, prec, out=o_np)   # restore o_np
        torch.cuda.synchronize()
    except Exception as exc:                       # extra information only: never fail the bench line over it
        e2e_pcm = {"error": str(exc)[:200]}
    # the e2e result must be the same numbers as the device-resident run (f64 host API converts on the device)
    chk = d_feat[:1000].double().cpu().numpy()
    assert np.allclose(o_np[:1000], chk, rtol=1e-6 if prec == "f32" else 1e-12, atol=0, equal_nan=True)

    if rank == 0:
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
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
            r0, F0, dt0 = cpu_port_rate(3600.0, cores, 0)                   # one of the workload's 1-hour files: ~15-20 s
            r1, F1, dt1 = cpu_port_rate(3600.0, cores, 1)
            cpu = {"value": r0, "unit": "frames/s", "cores": cores, "kind": "port",
                   "sample": f"one 1-hour synthetic mono 44.1 kHz file ({F0} frames), fft 2048 hop 512 Hann, C port of mp.rs structure "
                             f"(serial rstft + atan2 polar, analyzer on {cores} threads), {dt0:.1f} s",
                   "all_stages_parallel_value": r1}
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": max_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": prec, "data": "synthetic",
            "config": {"workload": f"config-5 shard: {files} x 1-hour 44.1 kHz mono files per GPU, fft 2048 hop 512 Hann, "
                                   f"14 features, one fused launch per step",
                       "files_per_gpu": files, "frames_per_step_per_gpu": frames, "l2": "inputs_larger_than_l2",
                       "sharding": "by file, no collective"},
            "audio_seconds_per_sec": value * HOP / SR,
            "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": e2e_files * FILE_SAMPLES * 8,
                    "d2h_bytes_per_step": e2e_frames * 14 * 8, "files_per_step": e2e_files, "steps": e2e_steps,
                    "host_dtype": "f64 (pinned)", "rank_cpu_affinity": numa},
            "e2e_pcm16": e2e_pcm,
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src, "kernel": "stft_kernel<double, 11, 0>" if prec == "f64" else "stft_kernel<float, 11, 1>",
                         "kernel_ms": k_ms, "algorithmic_bytes_per_launch": alg_bytes,
                         "note": "fused f64 path is co-limited by the FP64 pipe, issue slots and the L1/shared data pipe, not by HBM "
                                 "(DESIGN.md section 5); the spectra-writing kernel of the same path (rstft, spectra out) runs at ~73 % (profiles/r1s2_extra_configs.jsonl)"},
            "roofline_fp64": compute,
            "cpu_baseline": cpu,
            "clocks": clocks,
        }))
    if world > 1:
        dist.destroy_process_group()


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
    ap.add_argument("--direct-load", type=int, default=None, help="experiment: 1 = global->register input loads, 0 = TMA staging")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
