#!/usr/bin/env python
"""bench.py -- PCM samples/sec encoded by the B200 frame-encode path (BASELINE.json metric).

A "step" is one pass of the hot path (ingest -> analysis -> Rice search -> bit packing -> CRC) over one
batch: BASELINE config 2, one hour of 44.1 kHz / 16-bit stereo (38 760 frames of 4096), default encoder
config, synthetic noisy-sine PCM (flacenc_rs_b200.sigen.noisy_sine_pcm).

  value : inter-channel samples/s, whole job, inputs and outputs resident in HBM (fb200_encode_device),
          timed with CUDA events on the library's stream, max over ranks.
  e2e   : same metric through fb200_encode_interleaved with pinned HOST buffers, H2D and D2H inside the
          timed region.
  roofline     : dominant kernel of the step vs the measured HBM copy peak (MEASURED_PEAKS.json).
  cpu_baseline : the C oracle (a port of the reference's scalar path) on the host cores, bounded sample.

`--impl reference` times that CPU port with all host threads instead (the reference itself is a Rust
crate and cannot be built in this image).  N > 1: one process per GPU (torchrun), every rank encodes its
own one-hour stream (frames are independent; no collective on the data path), weak scaling.
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

CHANNELS, BPS, RATE, BLOCK = 2, 16, 44100, 4096
SECONDS = 3600
N_SAMPLES = SECONDS * RATE  # 158 760 000 per channel -> 38 760 frames (tail 3136)
WORKLOAD = "C2: 1 h 44.1 kHz 16-bit stereo, default config::Encoder, block 4096 (38760 frames)"


def make_pcm(rank: int, n: int) -> np.ndarray:
    """Synthetic noisy sine, a different seed per rank; the whole signal is generated, nothing is tiled."""
    from flacenc_rs_b200 import sigen
    return sigen.noisy_sine_pcm(n, CHANNELS, BPS, RATE, config_id=2 + 16 * rank)


def bind_to_gpu_numa_node(local_rank: int):
    """Pins this process to the CPUs NVML reports as local to its GPU, so the pinned staging buffers (first touch) and
    the copy threads sit on the GPU's NUMA node -- with one process per GPU the host links are otherwise shared badly.
    Returns (original affinity, description); a no-op when NVML or the topology is unavailable."""
    orig = os.sched_getaffinity(0)
    try:
        import pynvml
        import torch
        pynvml.nvmlInit()
        props = torch.cuda.get_device_properties(local_rank)
        handle = None
        uuid = getattr(props, "uuid", None)
        if uuid is not None:
            try:
                handle = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + str(uuid)).encode())
            except Exception:
                handle = None
        if handle is None:
            handle = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(handle, (ncpu + 63) // 64)
        cpus = {i for i in range(ncpu) if (words[i // 64] >> (i % 64)) & 1} & orig
        if cpus and cpus != orig:
            os.sched_setaffinity(0, cpus)
            return orig, f"{len(cpus)} of {len(orig)} CPUs (NVML affinity of the GPU)"
        return orig, "all CPUs (NVML reports no narrower affinity)"
    except Exception as e:  # noqa: BLE001
        return orig, f"unchanged ({type(e).__name__})"


class NvmlClockSampler:
    """The same readings through NVML calls in this process (a light thread, 5 ms period)."""

    def __init__(self, device: int):
        self.device = device
        self.sm, self.smax, self.reasons = [], [], set()
        self.stop_flag = threading.Event()
        self.thread = None

    def _run(self, handle):
        import pynvml
        bits = {"hw_slowdown": pynvml.nvmlClocksThrottleReasonHwSlowdown,
                "hw_thermal_slowdown": pynvml.nvmlClocksThrottleReasonHwThermalSlowdown,
                "sw_thermal_slowdown": pynvml.nvmlClocksThrottleReasonSwThermalSlowdown,
                "sw_power_cap": pynvml.nvmlClocksThrottleReasonSwPowerCap}
        while not self.stop_flag.is_set():
            try:
                self.sm.append(float(pynvml.nvmlDeviceGetClockInfo(handle, pynvml.NVML_CLOCK_SM)))
                self.smax.append(float(pynvml.nvmlDeviceGetMaxClockInfo(handle, pynvml.NVML_CLOCK_SM)))
                mask = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(handle)
                for name, bit in bits.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self.stop_flag.wait(0.005)

    def start(self):
        try:
            import pynvml
            import torch
            pynvml.nvmlInit()
            uuid = getattr(torch.cuda.get_device_properties(self.device), "uuid", None)
            handle = None
            if uuid is not None:
                try:
                    handle = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + str(uuid)).encode())
                except Exception:
                    handle = None
            if handle is None:
                handle = pynvml.nvmlDeviceGetHandleByIndex(self.device)
            self.thread = threading.Thread(target=self._run, args=(handle,), daemon=True)
            self.thread.start()
        except Exception:
            self.thread = None

    def stop(self) -> dict:
        if not self.thread:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["NVML unavailable"]}
        self.stop_flag.set()
        self.thread.join(timeout=2)
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None,
                "sm_max_mhz": max(self.smax) if self.smax else None, "samples": len(self.sm),
                "reasons": sorted(self.reasons), "source": "nvml"}


class OffSampler:
    def __init__(self, device: int):
        pass

    def start(self):
        pass

    def stop(self) -> dict:
        return {"sm_mhz": None, "sm_max_mhz": None, "samples": 0, "reasons": ["sampling off"]}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, device: int):
        self.device = device
        self.rows = []
        self.proc = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.device}", f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                smax.append(float(r[1]))
                for name, v in zip(names, r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def cpu_port_throughput(pcm: np.ndarray, threads: int, target_seconds: float):
    """Times the oracle (port of the reference's CPU path, par.rs-style frame workers) on a bounded sample:
    as many whole passes over `pcm` as fit in about `target_seconds` (at least one)."""
    from oracle import oracle as O
    cfg = O.default_config()
    frames = (len(pcm) + BLOCK - 1) // BLOCK  # incl. a shorter last frame
    sample = pcm
    probe = sample[: max(threads * 8, 64) * BLOCK]
    O.encode_frames(cfg, probe, CHANNELS, BPS, RATE, BLOCK, nthreads=threads)  # warm the thread pool / page in
    passes, total_s = 0, 0.0
    data, sizes = b"", []
    while passes == 0 or (total_s < target_seconds and passes < 64):
        t0 = time.perf_counter()
        data, sizes = O.encode_frames(cfg, sample, CHANNELS, BPS, RATE, BLOCK, nthreads=threads)
        total_s += time.perf_counter() - t0
        passes += 1
    return passes * len(sample) / total_s, frames, total_s, (data, sizes), passes


def compare_with_oracle(gpu_bytes: np.ndarray, gpu_sizes: np.ndarray, ref_bytes: bytes, ref_sizes: np.ndarray) -> dict:
    """Byte-compares the frames the timed GPU step produced with the oracle's frames of the same input (the second
    half of BASELINE.json's metric: stream size ratio GPU / reference, 1.0 when every decision matches)."""
    ref = np.frombuffer(ref_bytes, np.uint8)
    nf = min(len(gpu_sizes), len(ref_sizes))
    differing = abs(len(gpu_sizes) - len(ref_sizes))
    if len(gpu_bytes) == len(ref) and np.array_equal(gpu_sizes[:nf], ref_sizes[:nf]):
        if not np.array_equal(gpu_bytes, ref):
            neq = (gpu_bytes != ref).astype(np.uint8)
            starts = np.concatenate([[0], np.cumsum(ref_sizes[:nf], dtype=np.int64)[:-1]])
            differing += int(np.count_nonzero(np.add.reduceat(neq, starts)))
    else:
        go = np.concatenate([[0], np.cumsum(gpu_sizes, dtype=np.int64)])
        ro = np.concatenate([[0], np.cumsum(ref_sizes, dtype=np.int64)])
        for i in range(nf):
            if gpu_sizes[i] != ref_sizes[i] or not np.array_equal(gpu_bytes[go[i]:go[i + 1]], ref[ro[i]:ro[i + 1]]):
                differing += 1
    return {"frames_compared": int(max(len(gpu_sizes), len(ref_sizes))), "frames_differing": int(differing),
            "gpu_bytes": int(len(gpu_bytes)), "oracle_bytes": int(len(ref)),
            "stream_size_ratio_vs_oracle": len(gpu_bytes) / max(len(ref), 1)}


MODEL_OPS_PER_SAMPLE = 250.0  # minimal integer/FP lane-ops per inter-channel sample (stereo, default config)
NCU_METRICS = os.path.join(ROOT, "profiles", "ncu_metrics.json")  # written by tools/ncu_summary.py from the committed capture


def load_ncu_metrics() -> dict:
    """Per-kernel numbers of the committed `ncu --set full` capture (profiles/ncu_metrics.json, written by
    tools/ncu_summary.py): dram bytes, warp instructions and frames of the profiled launch.  Nothing is hard-coded here."""
    try:
        return json.load(open(NCU_METRICS))
    except Exception:
        return {}


def issue_roofline(value_per_gpu: float, clocks: dict, ncu: dict) -> dict:
    sm_mhz = float(clocks.get("sm_mhz") or 1965.0)
    peak = 148 * 4 * 32 * sm_mhz * 1e6  # SMs x schedulers x lanes x clock: lane-ops/s of one issue slot per scheduler
    out = {"bound": "issue", "model_ops_per_sample": MODEL_OPS_PER_SAMPLE, "peak_lane_ops_per_s": peak,
           "peak_samples_per_s": peak / MODEL_OPS_PER_SAMPLE, "frac": value_per_gpu * MODEL_OPS_PER_SAMPLE / peak,
           "note": "frac = per-GPU value x 250 ops / (148 SM x 4 x 32 lanes x SM clock); executed_frac uses the warp "
                   "instructions the kernels really issued in the committed ncu capture (profiles/ncu_metrics.json)"}
    k = ncu.get("kernels") or {}
    if k:
        per_frame = sum(v["warp_inst"] / max(v["frames"], 1) for v in k.values())
        executed = per_frame * 32.0 / BLOCK  # lane-ops per inter-channel sample actually issued
        out.update({"executed_ops_per_sample": executed, "executed_frac": value_per_gpu * executed / peak,
                    "warp_inst_per_frame": {n: v["warp_inst"] / max(v["frames"], 1) for n, v in k.items()},
                    "ncu_capture": ncu.get("tag")})
    return out


def copy_ceiling(torch, dist, world: int, device: int, in_bytes: int, out_bytes: int, reps: int = 6) -> dict:
    """The box's bare duplex copy ceiling at this N: every rank copies `in_bytes` host->device and `out_bytes`
    device->host from / to pinned memory on two streams at once, all ranks at the same time; nothing else runs.  The
    e2e step cannot be faster than this (it moves exactly these bytes)."""
    n_in, n_out = min(in_bytes, 256 << 20), min(out_bytes, 256 << 20)
    h_in = torch.empty(n_in, dtype=torch.uint8, pin_memory=True)
    h_out = torch.empty(n_out, dtype=torch.uint8, pin_memory=True)
    d_in = torch.empty(n_in, dtype=torch.uint8, device="cuda")
    d_out = torch.empty(n_out, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]

    def once():
        with torch.cuda.stream(s1):
            ev[0].record()
            d_in.copy_(h_in, non_blocking=True)
            ev[1].record()
        with torch.cuda.stream(s2):
            ev[2].record()
            h_out.copy_(d_out, non_blocking=True)
            ev[3].record()
        torch.cuda.synchronize()
        return ev[0].elapsed_time(ev[1]), ev[2].elapsed_time(ev[3])

    once()
    best_in = best_out = 1e9
    for _ in range(reps):
        if world > 1:
            dist.barrier()
        a, b = once()
        best_in, best_out = min(best_in, a), min(best_out, b)
    t = torch.tensor([best_in, best_out], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_in, ms_out = t.tolist()
    gbs_in, gbs_out = n_in / ms_in / 1e6, n_out / ms_out / 1e6  # per rank, under the load of all ranks
    floor_ms = max(in_bytes / (gbs_in * 1e6), out_bytes / (gbs_out * 1e6))
    return {"h2d_gbs_per_gpu": gbs_in, "d2h_gbs_per_gpu": gbs_out, "aggregate_gbs": (gbs_in + gbs_out) * world,
            "step_floor_ms": floor_ms,
            "how": f"pinned cudaMemcpyAsync of {n_in >> 20} MiB in + {n_out >> 20} MiB out on two streams at once, on all "
                   f"{world} ranks simultaneously, slowest rank, best of {reps}"}


def run_reference(args, rank: int, world: int) -> None:
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    # every step is one pass over the whole workload of the GPU arm (the full hour: 2-3 s of CPU wall on a 16-core
    # host), minus the short tail frame the frame-worker pool of the oracle binding does not take
    n = args.seconds * RATE
    pcm = make_pcm(0, n)
    vals = []
    frames = secs = 0
    for i in range(args.warmup + args.steps):
        v, frames, secs, _, _ = cpu_port_throughput(pcm, threads, target_seconds=0.0)
        if i >= args.warmup:
            vals.append(v)
    value = float(np.mean(vals))
    sample = f"all {frames} frames ({n / RATE:.1f} s of audio) of the workload per step, {secs:.2f} s CPU wall per pass"
    line = {
        "impl": "reference", "metric": "PCM inter-channel samples/sec encoded", "value": value, "unit": "samples/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1000.0 * n / value, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int32+f64", "data": "synthetic",
        "config": {"workload": WORKLOAD if args.seconds == SECONDS else f"{args.seconds} s slice of " + WORKLOAD,
                   "frames_per_step": frames, "sample": sample,
                   "note": "CPU port (C oracle) of the reference's scalar path with par.rs-style frame workers; "
                           "the Rust reference cannot be built in this image"},
        "cpu_baseline": {"value": value, "unit": "samples/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


CONFIG_CASES = [
    # name, channels, bps, rate, block, seconds, config kwargs (oracle-style)
    ("C1", "10 s CD stereo, default config (108 frames)", 2, 16, 44100, 4096, 10, {}),
    ("C3", "96 kHz / 24-bit stereo, block 4608, lpc_order 24 (reference maximum), 10 min = 12500 frames", 2, 24, 96000, 4608, 600,
     {"lpc_order": 24}),
    ("C4", "CD stereo 10 min, report/experimental.config.toml: use_direct_mse (covariance-method LPC) + Rectangle window",
     2, 16, 44100, 4096, 600, {"window_type": 0, "use_direct_mse": 1}),
    ("C4-irls2", "CD stereo 2 min, C4's config + mae_optimization_steps = 2 (IRLS-MAE refinement, 3 weighted solves per channel variant)",
     2, 16, 44100, 4096, 120, {"window_type": 0, "use_direct_mse": 1, "mae_optimization_steps": 2}),
    ("C5-slice", "48 kHz / 24-bit 8 channels, 2 min (1407 frames)", 8, 24, 48000, 4096, 120, {}),
    ("C2-ext-o4p4", "CD stereo 10 min, default config + the opt-in extensions ext_lpc_order_search = 4 (LPC orders 10, 8, 6, 4, 2 "
     "from one autocorrelation) and ext_lpc_precision_search = 4 (15..11 coefficient bits); not reference-compatible frames, "
     "compared with the oracle's statement of the extensions",
     2, 16, 44100, 4096, 600, {"ext_lpc_order_search": 4, "ext_lpc_precision_search": 4}),
    ("C2-bitcount", "CD stereo 10 min, default config with fixed.order_sel = BitCount (an exact Rice search per fixed order)",
     2, 16, 44100, 4096, 600, {"fixed_order_sel": 0}),
]


def make_cfg(kw):
    from flacenc_rs_b200.config import Encoder
    e = Encoder()
    if "lpc_order" in kw:
        e.subframe_coding.qlpc.lpc_order = kw["lpc_order"]
    if kw.get("window_type") == 0:
        e.subframe_coding.qlpc.window.type = "Rectangle"
    if kw.get("use_direct_mse"):
        e.subframe_coding.qlpc.use_direct_mse = True
    e.subframe_coding.qlpc.mae_optimization_steps = kw.get("mae_optimization_steps", 0)
    e.subframe_coding.qlpc.ext_order_search = kw.get("ext_lpc_order_search", 0)
    e.subframe_coding.qlpc.ext_precision_search = kw.get("ext_lpc_precision_search", 0)
    if kw.get("fixed_order_sel") == 0:
        from flacenc_rs_b200.config import OrderSel
        e.subframe_coding.fixed.order_sel = OrderSel.BitCount()
    return e.into_verified()


def other_configs(torch, device: int, threads: int) -> dict:
    """The other BASELINE.json configurations on one GPU: device-resident and end-to-end throughput, and every frame
    byte-compared with the oracle."""
    from flacenc_rs_b200 import sigen
    from flacenc_rs_b200.encoder import Context, pack_samples
    from oracle import oracle as O
    out = {}
    for key, desc, ch, bps, rate, block, secs, kw in CONFIG_CASES:
        n = secs * rate
        x = sigen.noisy_sine_pcm(n, ch, bps, rate, config_id=3)
        cb = (bps + 7) // 8
        packed = pack_samples(x, cb)
        n_frames = (n + block - 1) // block
        with Context(make_cfg(kw), ch, bps, rate, block, device=device) as ctx:
            cap = n_frames * ctx.max_frame_bytes()
            d_in = torch.from_numpy(packed).cuda()
            d_out = torch.empty(cap, dtype=torch.uint8, device="cuda")
            h_in = torch.empty(packed.nbytes, dtype=torch.uint8, pin_memory=True)
            h_in.numpy()[:] = packed
            h_out = torch.empty(cap, dtype=torch.uint8, pin_memory=True)
            sizes = np.zeros(n_frames, np.uint32)
            for _ in range(3):
                olen, _ = ctx.encode_device(d_in.data_ptr(), cb, n, d_out.data_ptr(), cap, 0, sizes)
            ms = []
            for _ in range(5):
                olen, _ = ctx.encode_device(d_in.data_ptr(), cb, n, d_out.data_ptr(), cap, 0, sizes)
                t = ctx.timing()
                ms.append(t.total_ms)
            kern = {"ingest": t.k_ingest_ms, "analyze": t.k_analyze_ms, "plan": t.k_rice_ms, "pack": t.k_pack_ms,
                    "fallback+scan": t.k_gather_ms}
            fused, fb = int(t.fused_frames), int(t.fallback_frames)
            e2e = []
            for i in range(6):
                got, hs, _ = ctx.encode_interleaved(h_in.numpy(), cb, n, 0, out=h_out.numpy())
                if i >= 2:
                    e2e.append(ctx.timing().total_ms)
            dev_bytes = d_out[:olen].cpu().numpy()
        ref, ref_sizes = O.encode_frames(O.default_config(**kw), x, ch, bps, rate, block, nthreads=threads)
        cmp_h = compare_with_oracle(got, hs, ref, ref_sizes)
        cmp_d = compare_with_oracle(dev_bytes, sizes, ref, ref_sizes)
        out[key] = {"workload": desc, "frames": n_frames, "device_samples_per_s": n / (float(np.mean(ms)) / 1e3),
                    "device_ms": float(np.mean(ms)), "e2e_samples_per_s": n / (float(np.mean(e2e)) / 1e3),
                    "e2e_ms": float(np.mean(e2e)), "kernel_ms": {a: round(b, 4) for a, b in kern.items()},
                    "fused_frames": fused, "fallback_frames": fb, "compression_ratio": olen / packed.nbytes,
                    "frames_differing": cmp_h["frames_differing"] + cmp_d["frames_differing"],
                    "bytes_equal_oracle": cmp_h["frames_differing"] + cmp_d["frames_differing"] == 0 and
                                          cmp_h["gpu_bytes"] == cmp_h["oracle_bytes"]}
        del d_in, d_out, h_in, h_out
    return out


def c5_chunk_list(torch, rank: int, world: int, device: int, threads: int, total_frames: int = 40960, chunk_frames: int = 1024):
    """BASELINE config 5 as a chunk list: 48 kHz / 24-bit / 8-channel chunks of `chunk_frames` frames generated ON the
    device, chunk c encoded by rank c mod world with input and output resident in HBM (10 h of this do not fit in host
    memory: SURVEY.md 8d).  Returns this rank's (device ms, frames, first-chunk parity)."""
    from flacenc_rs_b200.encoder import Context
    from oracle import oracle as O
    ch, bps, rate, block = 8, 24, 48000, 4096
    n_chunks = total_frames // chunk_frames
    mine = [c for c in range(n_chunks) if c % world == rank]
    n = chunk_frames * block
    scale = float(1 << (bps - 1))
    pcms = []
    for c in mine:
        g = torch.Generator(device="cuda")
        g.manual_seed(0xC5000 + c)
        t = torch.arange(c * n, (c + 1) * n, device="cuda", dtype=torch.float64).unsqueeze(1)
        k = torch.arange(ch, device="cuda", dtype=torch.float64).unsqueeze(0)
        ramp = 0.75 + 0.25 * torch.sin(2.0 * np.pi * 0.1 * t / rate + 0.5 * k)
        sine = 0.8 * ramp * torch.sin(2.0 * np.pi * 440.0 * t / rate + 0.37 * k)
        noise = 0.4 * (torch.rand((n, ch), generator=g, device="cuda", dtype=torch.float64) - 0.5)
        x = torch.clamp(torch.round((sine + noise) * scale), -scale, scale - 1).to(torch.int32)
        pcms.append(x.view(torch.uint8).reshape(-1, 4)[:, :3].contiguous().reshape(-1))  # packed little-endian 24-bit
        del t, ramp, sine, noise, x
    ms, frames, parity = 0.0, 0, None
    with Context(make_cfg({}), ch, bps, rate, block, device=device) as ctx:
        cap = chunk_frames * ctx.max_frame_bytes()
        d_out = torch.empty(cap, dtype=torch.uint8, device="cuda")
        sizes = np.zeros(chunk_frames, np.uint32)
        if pcms:
            ctx.encode_device(pcms[0].data_ptr(), 3, n, d_out.data_ptr(), cap, mine[0] * chunk_frames, sizes)  # warm-up
        torch.cuda.synchronize()
        for c, pcm in zip(mine, pcms):
            olen, _ = ctx.encode_device(pcm.data_ptr(), 3, n, d_out.data_ptr(), cap, c * chunk_frames, sizes)
            ms += ctx.timing().total_ms
            frames += chunk_frames
            if rank == 0 and parity is None:  # the first 128 frames of the first chunk against the oracle
                k = 128
                raw = pcm[: k * block * ch * 3].cpu().numpy().reshape(-1, 3)
                x = (raw[:, 0].astype(np.int32) | (raw[:, 1].astype(np.int32) << 8) | (raw[:, 2].astype(np.int8).astype(np.int32) << 16))
                ref, ref_sizes = O.encode_frames(O.default_config(), x.reshape(-1, ch), ch, bps, rate, block,
                                                 first_frame_number=c * chunk_frames, nthreads=threads)
                got = d_out[: int(sizes[:k].sum())].cpu().numpy()
                parity = compare_with_oracle(got, sizes[:k].copy(), ref, ref_sizes)
    return ms, frames, parity


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--seconds", type=int, default=SECONDS, help="audio seconds per rank (default: the 1 h workload)")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the CPU legs (oracle baseline, byte compare, configs)")
    ap.add_argument("--no-extras", action="store_true", help="only the two timed regions (value, e2e)")
    ap.add_argument("--clock-sampler", default="auto", choices=["auto", "smi", "nvml", "off"],
                    help="how SM clocks / throttle reasons are sampled during the timed regions")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    from flacenc_rs_b200 import _ffi
    from flacenc_rs_b200.config import Encoder
    from flacenc_rs_b200.encoder import Context, encode_interleaved_sharded, pack_samples

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; flacenc_rs_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    cpu_group = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        cpu_group = dist.new_group(backend="gloo")  # waits that must not occupy the GPUs (rank 0's single-process legs)

    orig_affinity, affinity_note = bind_to_gpu_numa_node(local_rank)
    n = args.seconds * RATE
    n_frames = (n + BLOCK - 1) // BLOCK
    pcm_i32 = make_pcm(rank, n)
    packed = pack_samples(pcm_i32, 2)  # packed LE16 interleaved: what Fill::fill_le_bytes receives
    in_bytes = packed.nbytes

    ctx = Context(Encoder().into_verified(), CHANNELS, BPS, RATE, BLOCK, device=local_rank)
    cap = n_frames * ctx.max_frame_bytes()
    # device-resident buffers (value) and pinned host buffers (e2e)
    d_in = torch.from_numpy(packed).cuda()
    d_out = torch.empty(cap, dtype=torch.uint8, device="cuda")
    h_in_t = torch.empty(in_bytes, dtype=torch.uint8, pin_memory=True)
    h_in = h_in_t.numpy()
    h_in[:] = packed
    h_out_t = torch.empty(cap, dtype=torch.uint8, pin_memory=True)
    h_out = h_out_t.numpy()
    sizes = np.zeros(n_frames, np.uint32)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def cpu_barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier(group=cpu_group)

    def step_device():
        olen, _ = ctx.encode_device(d_in.data_ptr(), 2, n, d_out.data_ptr(), cap, 0, sizes)
        return olen, ctx.timing()

    host_sizes = [None]

    def step_host():
        got, hs, _ = ctx.encode_interleaved(h_in, 2, n, 0, out=h_out)
        host_sizes[0] = hs
        return len(got), ctx.timing()

    # ---- warm-up of the device-resident path (the host path is warmed right before its own timed region)
    for _ in range(args.warmup):
        out_len, _ = step_device()
    step_host()

    # ---- timed: device-resident (value).  Inputs (635 MB) and outputs (567 MB) are far larger than the 126 MB L2, so
    # every step streams from HBM.
    kind = args.clock_sampler
    if kind == "auto":  # NVML in-process gives a sample every 20 ms; the nvidia-smi loop (100 ms) is the fallback
        try:
            import pynvml  # noqa: F401
            kind = "nvml"
        except Exception:
            kind = "smi"
    sampler = {"smi": ClockSampler, "nvml": NvmlClockSampler, "off": OffSampler}[kind](local_rank)
    barrier()
    sampler.start()
    # "analyze" = K1 (windowed autocorrelation, Levinson, quantiser, fixed-order estimate), "plan" = KA (Rice search,
    # decisions, frame plan), "pack" = KP (bit packing + CRC + store at the final offset), "fallback+scan" = the generic
    # kernels over the frames KA handed back (normally none) + the size scan; "ingest" only exists when the input is
    # not 16-bit stereo (the packed PCM is read in place otherwise)
    dev_ms, kern = 0.0, {"ingest": 0.0, "analyze": 0.0, "plan": 0.0, "pack": 0.0, "fallback+scan": 0.0}
    fused_frames = fallback_frames = 0
    launches = 0
    w0 = time.perf_counter()
    for _ in range(args.steps):
        out_len, t = step_device()
        dev_ms += t.total_ms
        kern["ingest"] += t.k_ingest_ms
        kern["analyze"] += t.k_analyze_ms
        kern["plan"] += t.k_rice_ms
        kern["pack"] += t.k_pack_ms
        kern["fallback+scan"] += t.k_gather_ms
        launches += t.launches
        fused_frames, fallback_frames = t.fused_frames, t.fallback_frames
    barrier()
    wall_dev = time.perf_counter() - w0

    # ---- timed: end to end with host buffers (e2e): the library's own CUDA events bracket every H2D copy, kernel and
    # D2H copy of the call; the wall clock around the loop is reported next to it
    for _ in range(args.warmup):
        step_host()
    barrier()
    e2e_ms = 0.0
    h2d_ms = d2h_ms = 0.0
    e2e_steps = []
    w1 = time.perf_counter()
    for _ in range(args.steps):
        out_len_h, t = step_host()
        e2e_ms += t.total_ms
        e2e_steps.append(round(t.total_ms, 3))
        h2d_ms += t.h2d_ms
        d2h_ms += t.d2h_ms
    barrier()
    wall_e2e = time.perf_counter() - w1
    clocks = sampler.stop()
    assert out_len_h == out_len

    # max over ranks of the device time; aggregate = all ranks' samples / that time
    times = torch.tensor([dev_ms, e2e_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    dev_ms_max, e2e_ms_max = times.tolist()
    total_samples = float(n) * world * args.steps
    value = total_samples / (dev_ms_max / 1000.0)
    e2e_value = total_samples / (e2e_ms_max / 1000.0)

    # ---- the box's bare duplex copy ceiling at this N (all ranks at once): the floor of the e2e step
    extras = not args.no_extras
    ceiling = copy_ceiling(torch, dist, world, local_rank, in_bytes, int(out_len)) if extras else None

    # ---- BASELINE config 5 as a device-generated chunk list, chunk c on rank c mod N (strong scaling over the ranks)
    c5 = None
    if extras:
        threads_all = os.cpu_count() or 1
        barrier()
        c5_ms, c5_frames, c5_parity = c5_chunk_list(torch, rank, world, local_rank, threads_all)
        t5 = torch.tensor([c5_ms], dtype=torch.float64, device="cuda")
        f5 = torch.tensor([float(c5_frames)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t5, op=dist.ReduceOp.MAX)
            dist.all_reduce(f5, op=dist.ReduceOp.SUM)
        c5 = {"workload": "C5 chunk list: 48 kHz / 24-bit / 8 ch, 40 chunks of 1024 frames generated on the device, chunk c on "
                          "rank c mod N, input and output resident in HBM",
              "frames": int(f5.item()), "ms": t5.item(), "scaling": "strong",
              "samples_per_s": f5.item() * BLOCK / (t5.item() / 1e3), "pcm_samples_per_s": f5.item() * BLOCK * 8 / (t5.item() / 1e3),
              "first_128_frames_vs_oracle": c5_parity}

    # ---- one stream sharded by frame range over the N devices of the box, driven by ONE process (rank 0): the multi-GPU
    # path the north star names (src/par.rs:355-449 replaced by fb200_encode_interleaved_sharded).  The other ranks wait on
    # a CPU barrier, so their GPUs are free.  Strong scaling: the same stream as the N = 1 e2e number.
    sharded = None
    cpu_barrier()
    if extras and rank == 0:
        devs = list(range(world))
        ctxs = [ctx] + [Context(Encoder().into_verified(), CHANNELS, BPS, RATE, BLOCK, device=d) for d in devs[1:]]
        h_out2_t = torch.empty(cap, dtype=torch.uint8, pin_memory=True)
        h_out2 = h_out2_t.numpy()
        for _ in range(max(2, args.warmup)):
            got2, sizes2 = encode_interleaved_sharded(ctxs, h_in, 2, n, 0, out=h_out2)
        sh_ms, sh_steps = 0.0, []
        ws = time.perf_counter()
        for _ in range(args.steps):
            got2, sizes2 = encode_interleaved_sharded(ctxs, h_in, 2, n, 0, out=h_out2)
            tm = ctx.timing()
            sh_ms += tm.total_ms
            sh_steps.append(round(tm.total_ms, 3))
        wall_sh = time.perf_counter() - ws
        same = len(got2) == out_len_h and np.array_equal(got2, h_out[:out_len_h]) and np.array_equal(sizes2, host_sizes[0])
        # the copies of ONE stream spread over N devices cannot beat the box's duplex ceiling at this N either
        sh_floor = max(in_bytes / (world * ceiling["h2d_gbs_per_gpu"] * 1e6), out_len / (world * ceiling["d2h_gbs_per_gpu"] * 1e6))
        sharded = {"value": float(n) * args.steps / (sh_ms / 1e3), "unit": "samples/s", "n_devices": world, "scaling": "strong",
                   "ms_per_step": sh_ms / args.steps, "ms_steps": sh_steps, "wall_ms_per_step": 1e3 * wall_sh / args.steps,
                   "copy_floor_ms": sh_floor, "frac_of_copy_ceiling": sh_floor / (sh_ms / args.steps),
                   "bytes_equal_to_single_gpu_result": bool(same),
                   "api": "fb200_encode_interleaved_sharded: one process, one context per device, chunk c on device c mod N, "
                          "every chunk copied to its final offset in the caller's pinned buffer; ms_per_step = longest "
                          "per-device span (CUDA events), wall clock next to it"}
        for c in ctxs[1:]:
            c.close()
        del h_out2_t
    cpu_barrier()

    os.sched_setaffinity(0, orig_affinity)  # the CPU legs below use every host core again
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        ncu = load_ncu_metrics()
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_kind = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
        dom = max(kern, key=kern.get)
        alg_bytes = float(in_bytes + out_len)  # algorithmic bytes of one step: packed PCM in + frame bytes out
        dom_s = kern[dom] / args.steps / 1000.0
        achieved = alg_bytes / dom_s / 1e9
        whole = alg_bytes / (dev_ms / args.steps / 1000.0) / 1e9
        ncu_dom = (ncu.get("kernels") or {}).get(dom)
        traffic = ncu_dom["dram_bytes"] / max(ncu_dom["frames"], 1) * n_frames if ncu_dom else None
        line = {
            "metric": "PCM inter-channel samples/sec encoded", "value": value, "unit": "samples/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int32+f64", "data": "synthetic",
            "config": {"workload": WORKLOAD if args.seconds == SECONDS else f"{args.seconds} s slice of " + WORKLOAD,
                       "pcm_samples_per_s": value * CHANNELS, "frames_per_step": n_frames,
                       "compression_ratio": out_len / in_bytes,
                       "stream_size_ratio_vs_oracle": None, "frames_differing": None, "frames_compared": 0,
                       "fused_frames": int(fused_frames), "fallback_frames": int(fallback_frames),
                       "l2": "inputs (635 MB/step) and outputs (567 MB/step) exceed the 126 MB L2; no flush needed",
                       "cpu_affinity": affinity_note,
                       "timing": "CUDA events on the library stream (fb200_last_timing), max over ranks"},
            "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": int(in_bytes),
                    "d2h_bytes_per_step": int(out_len + 4 * n_frames + 16),
                    "h2d_ms_per_step": h2d_ms / args.steps, "d2h_ms_per_step": d2h_ms / args.steps,
                    "ms_per_step": e2e_ms_max / args.steps, "ms_steps_rank0": e2e_steps,
                    "copy_ceiling": ceiling,
                    "frac_of_copy_ceiling": (ceiling["step_floor_ms"] / (e2e_ms_max / args.steps)) if ceiling else None,
                    "api": "fb200_encode_interleaved, pinned host buffers; chunks pipelined over H2D / compute / D2H "
                           "streams (h2d_ms / d2h_ms are summed copy times and overlap the kernels)"},
            "gpu_launches": int(launches),
            "kernel_ms_per_step": {k: v / args.steps for k, v in kern.items()},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic,
                         "traffic_note": "dram bytes read+written by the dominant kernel per step: bytes per frame of the "
                                         "committed ncu capture (profiles/ncu_metrics.json) x frames of this step",
                         "kernel": dom, "peak_kind": peak_kind,
                         "algorithmic_bytes_per_step": alg_bytes,
                         "whole_step_achieved_gbs": whole, "whole_step_frac": whole / peak,
                         "note": "the path is instruction-issue bound (hundreds of integer/FP ops per 6.2 algorithmic bytes): "
                                 "the HBM fraction is low by construction, see issue_roofline and DESIGN.md"},
            # instruction-issue bound of the analysis (the binding resource): minimal op model of the
            # reference algorithm vs the INT32/FP32 lane-op rate of ONE GPU at the sampled SM clock
            "issue_roofline": issue_roofline(value / world, clocks, ncu),
            "clocks": clocks,
            "wall_s_device_loop": wall_dev, "wall_s_e2e_loop": wall_e2e,
        }
        if sharded:
            line["sharded"] = sharded
        if c5:
            line["c5_chunk_list"] = c5
        if extras and not args.no_cpu_baseline and world == 1:
            threads = os.cpu_count() or 1
            line["configs"] = other_configs(torch, local_rank, threads)
            # ---- stream level: fLaC + STREAMINFO + MD5 around the frame path.  MD5 of the PCM is
            # sequential host work per stream (src/source.rs:406-429); md5_floor = this library's MD5 alone over the same
            # bytes on one core.  The batch call hashes every stream on its own thread.
            from flacenc_rs_b200.encoder import encode_streams_with_fixed_block_size, encode_with_fixed_block_size
            from flacenc_rs_b200.source import MemSource
            ns = min(n, 600 * RATE)
            src = MemSource.from_samples(pcm_i32[:ns], CHANNELS, BPS, RATE)
            encode_with_fixed_block_size(Encoder().into_verified(), src, BLOCK, devices=[local_rank])  # warm (contexts are kept)
            dts = []
            for _ in range(3):
                t0 = time.perf_counter()
                stream = encode_with_fixed_block_size(Encoder().into_verified(), src, BLOCK, devices=[local_rank])
                dts.append(time.perf_counter() - t0)
            dt = min(dts)
            dig = np.zeros(16, np.uint8)
            md5_in = packed[: ns * CHANNELS * 2]
            t0 = time.perf_counter()
            _ffi.lib().fb200_md5(md5_in.ctypes.data, md5_in.nbytes, dig.ctypes.data)
            md5_s = time.perf_counter() - t0
            from oracle import oracle as O
            ref_stream = O.encode_stream(O.default_config(), pcm_i32[:ns], CHANNELS, BPS, RATE, BLOCK, nthreads=threads)
            line["stream_api"] = {"value": ns / dt, "unit": "samples/s", "seconds_of_audio": ns / RATE, "wall_s": dt,
                                  "stream_bytes": len(stream), "equals_oracle_stream": bool(stream.write() == ref_stream),
                                  "md5_floor": {"value": ns / md5_s, "unit": "samples/s", "gb_per_s": md5_in.nbytes / md5_s / 1e9,
                                                "wall_s": md5_s},
                                  "frac_of_md5_floor": md5_s / dt,
                                  "note": "encode_with_fixed_block_size (fb200_encode_stream) on pageable int32 samples, whole "
                                          "call incl. the host range check; MD5 runs on its own thread next to the device work"}
            k = 16
            seg = ns // 4  # sixteen 2.5-minute streams
            srcs = [MemSource.from_samples(pcm_i32[(i % 4) * seg:(i % 4 + 1) * seg], CHANNELS, BPS, RATE) for i in range(k)]
            encode_streams_with_fixed_block_size(Encoder().into_verified(), srcs[:2], BLOCK, devices=[local_rank])
            t0 = time.perf_counter()
            outs = encode_streams_with_fixed_block_size(Encoder().into_verified(), srcs, BLOCK, devices=[local_rank])
            dtb = time.perf_counter() - t0
            line["stream_batch"] = {"value": k * seg / dtb, "unit": "samples/s", "streams": k, "seconds_of_audio_each": seg / RATE,
                                    "wall_s": dtb, "host_threads": threads, "all_equal": bool(all(
                                        outs[i].write() == outs[i % 4].write() for i in range(k))),
                                    "note": "fb200_encode_streams: one MD5 thread per stream (bounded by the host's cores)"}
            _ffi.lib().fb200_pool_clear()
        if not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            # (timed as the CPU baseline at N = 1 only; at N > 1 a single pass serves the byte comparison of rank 0's stream)
            v, frames, secs, (ref_bytes, ref_sizes), passes = cpu_port_throughput(pcm_i32, threads,
                                                                                  target_seconds=12.0 if world == 1 else 0.0)
            if world == 1:
                line["cpu_baseline"] = {"value": v, "unit": "samples/s", "cores": threads, "kind": "port",
                                        "sample": f"{passes} passes over the same {frames} frames ({n / RATE:.1f} s of audio) "
                                                  f"the GPU step encodes, {secs:.2f} s wall, C oracle with {threads} frame workers"}
            # parity on the measured workload: the bytes the last timed e2e step left in the pinned host buffer (and
            # the device-resident step's bytes) against the oracle's frames of the same PCM
            cmp_host = compare_with_oracle(h_out[:out_len_h], host_sizes[0], ref_bytes, ref_sizes)
            cmp_dev = compare_with_oracle(d_out[:out_len].cpu().numpy(), sizes, ref_bytes, ref_sizes)
            line["config"].update({"stream_size_ratio_vs_oracle": cmp_host["stream_size_ratio_vs_oracle"],
                                   "frames_differing": cmp_host["frames_differing"] + cmp_dev["frames_differing"],
                                   "frames_compared": cmp_host["frames_compared"],
                                   "parity": {"e2e_step": cmp_host, "device_step": cmp_dev,
                                              "note": "every frame of the timed workload byte-compared with the C oracle"}})
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        cpu_barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
