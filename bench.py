#!/usr/bin/env python
"""bench.py - headline benchmark of the B200 audio feature path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Metric (BASELINE.json): featurized clips/s, augmentation + features, on BASELINE.json
configs[1]: MFCC-40 (n_fft 400, hop 160, 40 mels) + background-noise mixing at a target SNR +
RIR reverb, batch 1024 synthetic 1.5 s 16 kHz clips PER GPU (weak scaling; the batch is sharded
by clip index, no collective on the feature path).  One "step" = one pass of the hot path over
one batch.  Prints ONE JSON line (rank 0).

  value        whole-job clips/s with the clips already resident in HBM (CUDA events, max over ranks)
  e2e          same metric through the public API with HOST buffers: pinned host clips + draws
               -> H2D -> wwf_featurize -> D2H of the features, all inside the timed region
  roofline     dominant kernel: algorithmic bytes / its CUDA-event launch time vs measured HBM peak
  cpu_baseline the oracle (torchaudio CPU, the reference's arithmetic) on a bounded sample, rank 0, N=1

--impl reference times that torchaudio CPU path alone, on all host threads.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

# ---- workload: BASELINE.json configs[1] -----------------------------------------------------
SR, N_FFT, HOP, N_MELS, N_MFCC = 16000, 400, 160, 40, 40
B_PER_GPU, N_SAMPLES = 1024, 24000
N_NOISE, NOISE_LEN, N_RIR, RIR_LEN = 256, 24000, 64, 8000
RING = 4                       # distinct input batches cycled through: 4 x 98.3 MB > 126 MB L2
T_FRAMES = N_SAMPLES // HOP + 1
# algorithmic bytes per clip: clip + noise segment + RIR in, features out
BYTES_STEP = 4 * (N_SAMPLES + N_SAMPLES + RIR_LEN) + 4 * N_MFCC * T_FRAMES          # 248 160
BYTES_CONV = 4 * (N_SAMPLES + RIR_LEN + N_SAMPLES)                                  # x, h in; y out
BYTES_FEAT = 4 * (N_SAMPLES + N_SAMPLES) + 4 * N_MFCC * T_FRAMES                    # y, noise in; features out
# large-batch path: the frames kernel reads y + noise and writes the dB tile; the block epilogue reads it, writes features
BYTES_FRAMES = 4 * (N_SAMPLES + N_SAMPLES) + 4 * N_MELS * T_FRAMES
BYTES_EPILOGUE = 4 * N_MELS * T_FRAMES + 4 * N_MFCC * T_FRAMES
KERNEL_BYTES = {"conv_kernel": BYTES_CONV, "feat_kernel": BYTES_FEAT, "feat_frames_kernel": BYTES_FRAMES,
                "feat_epilogue_block_kernel": BYTES_EPILOGUE, "feat_epilogue_mma_kernel": BYTES_EPILOGUE, "feat_prep_kernel": 0}
METRIC = "featurized clips/sec (1.5s@16kHz, aug+log-mel+DCT: configs[1] MFCC-40 + noise@SNR + RIR)"
WORKLOAD = ("configs[1]: MFCC-40 (n_fft 400, hop 160, 40 mels) + noise@SNR U[5,20] dB + RIR reverb "
            "(8000 taps), batch 1024 x 1.5 s @ 16 kHz per GPU")


def synth(seed: int, B: int):
    g = torch.Generator().manual_seed(seed)
    wav = 0.1 * torch.randn(B, N_SAMPLES, generator=g)
    draws = dict(rir_idx=torch.randint(0, N_RIR, (B,), generator=g, dtype=torch.int32),
                 noise_idx=torch.randint(0, N_NOISE, (B,), generator=g, dtype=torch.int32),
                 noise_off=torch.randint(0, NOISE_LEN, (B,), generator=g),
                 snr_db=5.0 + 15.0 * torch.rand(B, generator=g))
    return wav, draws


def synth_banks():
    g = torch.Generator().manual_seed(1234)
    noise = [0.05 * torch.randn(NOISE_LEN, generator=g) for _ in range(N_NOISE)]
    t = torch.arange(RIR_LEN, dtype=torch.float32)
    rirs = [torch.randn(RIR_LEN, generator=g) * torch.exp(-t / 1000.0) for _ in range(N_RIR)]
    return noise, rirs


def ncu_traffic(kernel: str):
    """DRAM bytes per launch of `kernel` from the committed ncu capture (profiles/r01_traffic.json)."""
    try:
        with open(os.path.join(ROOT, "profiles", "r01_traffic.json")) as f:
            return int(json.load(f)[kernel])
    except Exception:
        return None


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ---- clocks ---------------------------------------------------------------------------------
class ClockSampler:
    """Polls NVML for SM clock and throttle reasons while the timed loops run."""
    BITS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
            0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x100: "display_clock_setting"}

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz, self.ok = [], set(), None, False
        self._stop = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception:
            self.ok = False
        self.th = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self._stop.is_set():
            try:
                self.samples.append(int(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)))
                r = int(self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                for bit, name in self.BITS.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        if self.ok:
            self.th.start()

    def stop(self):
        self._stop.set()
        if self.ok and self.th.is_alive():
            self.th.join(timeout=1.0)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


# ---- reference arm / CPU baseline ----------------------------------------------------------
def cpu_reference(n_clips: int, reps: int, warm: int):
    """Times the oracle (torchaudio CPU) on n_clips clips of the same workload; returns clips/s."""
    from oracle import ta_oracle as tao
    torch.set_num_threads(os.cpu_count() or 1)
    noise, rirs = synth_banks()
    wav, d = synth(7, n_clips)
    kw = dict(rirs=rirs, rir_idx=d["rir_idx"], noise_bank=noise, noise_idx=d["noise_idx"], noise_off=d["noise_off"],
              snr_db=d["snr_db"], sample_rate=SR, feature_type="mfcc", n_mels=N_MELS, n_mfcc=N_MFCC, n_fft=N_FFT,
              hop_length=HOP)
    for _ in range(warm):
        tao.pipeline(wav, **kw)
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        tao.pipeline(wav, **kw)
        ts.append(time.perf_counter() - t0)
    return n_clips * len(ts) / sum(ts), sum(ts) / len(ts), torch.get_num_threads()


def cpu_per_clip_loop(n_clips: int = 24):
    """The reference's real usage pattern (one call chain per clip, src/evaluation/evaluator.py:202-205,
    Dataset.__getitem__): clips/s of the oracle called clip by clip."""
    from oracle import ta_oracle as tao
    noise, rirs = synth_banks()
    wav, d = synth(8, n_clips)
    kw = dict(sample_rate=SR, feature_type="mfcc", n_mels=N_MELS, n_mfcc=N_MFCC, n_fft=N_FFT, hop_length=HOP)
    def one(i):
        tao.pipeline(wav[i:i + 1], rirs=rirs, rir_idx=d["rir_idx"][i:i + 1], noise_bank=noise, noise_idx=d["noise_idx"][i:i + 1],
                     noise_off=d["noise_off"][i:i + 1], snr_db=d["snr_db"][i:i + 1], **kw)
    one(0)
    t0 = time.perf_counter()
    for i in range(n_clips):
        one(i)
    return n_clips / (time.perf_counter() - t0)


def cpu_shape_augs(n_stretch: int = 32, n_pitch: int = 8):
    """torchaudio CPU (oracle/ta_oracle.py) on a bounded sample: time-stretch and F.pitch_shift, clips/s."""
    from oracle import ta_oracle as tao
    g = torch.Generator().manual_seed(3)
    x = 0.1 * torch.randn(max(n_stretch, n_pitch), N_SAMPLES, generator=g)
    rates = 0.8 + 0.4 * torch.rand(n_stretch, generator=g, dtype=torch.float64)
    t0 = time.perf_counter()
    tao.time_stretch(x[:n_stretch], rates)
    t1 = time.perf_counter()
    tao.pitch_shift(x[:n_pitch], torch.tensor([-2, -1, 1, 2] * (n_pitch // 4), dtype=torch.int32), 16000)
    t2 = time.perf_counter()
    return {"time_stretch": n_stretch / (t1 - t0), "pitch_shift": n_pitch / (t2 - t1), "unit": "clips/s",
            "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{n_stretch} clips time-stretch, {n_pitch} clips F.pitch_shift (1.5 s each), torchaudio CPU"}


def run_reference(args, rank: int):
    if rank != 0:
        return
    n_clips = 256 if args.steps <= 60 else (128 if args.steps <= 150 else 64)   # keep the whole run to a few minutes
    cps, sec, cores = cpu_reference(n_clips, reps=args.steps, warm=max(1, min(args.warmup, 3)))
    line = {"impl": "reference", "metric": METRIC, "value": cps, "unit": "clips/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "sample_per_step": f"{n_clips} clips of the workload, one batched "
                       "torchaudio call chain (fftconvolve per RIR group, add_noise, MFCC) on CPU"},
            "cpu_baseline": {"value": cps, "unit": "clips/s", "cores": cores, "kind": "port",
                             "sample": f"{n_clips} clips/step x {args.steps} steps; oracle/ta_oracle.py = the "
                                       "reference's torchaudio CPU arithmetic (its src/data module is absent upstream)"},
            "e2e": {"value": cps, "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---- our arm -----------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch.distributed as dist
    import wakeword_trainer_home_b200 as w

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the feature path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    from wakeword_trainer_home_b200.sharding import bind_to_gpu_numa, numa_local
    numa_bound = bind_to_gpu_numa(local_rank) if world > 1 else False   # NUMA-local pinned buffers per rank
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local_rank])
        torch.cuda.synchronize(dev)

    def max_over_ranks(ms: float) -> float:
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # plan + banks (replicated per GPU), device-resident ring of batches for this rank's shard
    B = B_PER_GPU
    plan = w.FeaturePlan(SR, "mfcc", N_MELS, N_MFCC, N_FFT, HOP, dev)
    noise, rirs = synth_banks()
    plan.register_noise(noise)
    plan.register_rirs(rirs)
    host = [synth(100 * rank + i, B) for i in range(RING)]
    # pinned host buffers are allocated (and first touched) on the cores next to this GPU's PCIe root; a
    # single-GPU run gets its full CPU affinity back afterwards so the CPU baseline still uses every core
    with numa_local(local_rank) as nl:
        pinned_wav = [h[0].pin_memory() for h in host]
        pinned_draws = [{k: v.pin_memory() for k, v in h[1].items()} for h in host]
        sf = w.StreamedFeaturizer(plan, B, N_SAMPLES, depth=2, copy_back=True)
        pinned_pcm = [(h.clamp(-1, 1) * 32767).to(torch.int16).pin_memory() for h in pinned_wav]
        sf16 = w.StreamedFeaturizer(plan, B, N_SAMPLES, depth=2, copy_back=True, pcm16=True)
    numa_bound = numa_bound or nl.bound
    dev_wav = [h.to(dev) for h in pinned_wav]
    dev_aug = [w.AugParams(**d).to(dev) for d in pinned_draws]
    out = torch.empty(B, 1, N_MFCC, T_FRAMES, dtype=torch.float32, device=dev)
    stream = torch.cuda.current_stream(dev)

    def step(i):
        plan.featurize(dev_wav[i % RING], dev_aug[i % RING], out=out)

    for i in range(args.warmup):
        step(i)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()

    # ---- device-resident timing: exactly K steps between two events ----
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n0 = w.launch_count()
    barrier()
    e0.record(stream)
    for i in range(args.steps):
        step(i)
    e1.record(stream)
    barrier()
    launches = w.launch_count() - n0
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    ms_step = ms_total / args.steps
    value = world * B * args.steps / (ms_total * 1e-3)

    # ---- per-kernel launch times: the same K steps with the library's event hook on ----
    plan.profile(True)
    for i in range(args.steps):
        step(i)
    kernel_ms, _, n_split = plan.profile_read_kernels()   # waits for the events; averages per call, per kernel
    plan.profile(False)
    conv_ms = kernel_ms.get("conv_kernel", 0.0)
    feat_ms = sum(v for k, v in kernel_ms.items() if k != "conv_kernel")

    # ---- end to end through the public API with HOST buffers (pinned), copies inside the timed
    #      region: upload of clips + draws, featurize, download of the features, triple-streamed ----
    host_aug = [w.AugParams(**d) for d in pinned_draws]
    for i in range(4):
        sf.submit(pinned_wav[i % RING], host_aug[i % RING])
    sf.synchronize()
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e2.record(sf.s_in)
    for i in range(args.steps):
        sf.submit(pinned_wav[i % RING], host_aug[i % RING])
    e3.record(sf.s_out)
    sf.synchronize()
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    e2e_ms = max_over_ranks(max(e2.elapsed_time(e3), 0.0))
    e2e_value = world * B * args.steps / (e2e_ms * 1e-3)
    h2d = pinned_wav[0].numel() * 4 + host_aug[0].nbytes()
    d2h = sf.h_out[0].numel() * sf.h_out[0].element_size()
    # what the link gives: a bare pinned -> device copy of one clip batch, best of 5 (the e2e path's ceiling)
    link_ms = 1e9
    for _ in range(5):
        ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ea.record(stream)
        dev_wav[0].copy_(pinned_wav[0], non_blocking=True)
        eb.record(stream)
        eb.synchronize()
        link_ms = min(link_ms, ea.elapsed_time(eb))
    link_gbs = pinned_wav[0].numel() * 4 / (link_ms * 1e-3) / 1e9
    e2e_h2d_gbs = h2d / (e2e_ms / args.steps * 1e-3) / 1e9

    # ---- supplementary: the same host-fed pipeline with int16 PCM host buffers (half the H2D bytes) ----
    for i in range(4):
        sf16.submit(pinned_pcm[i % RING], host_aug[i % RING])
    sf16.synchronize()
    barrier()
    e6, e7 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e6.record(sf16.s_in)
    for i in range(args.steps):
        sf16.submit(pinned_pcm[i % RING], host_aug[i % RING])
    e7.record(sf16.s_out)
    sf16.synchronize()
    barrier()
    pcm_ms = max_over_ranks(max(e6.elapsed_time(e7), 0.0))
    pcm_value = world * B * args.steps / (pcm_ms * 1e-3)

    # ---- supplementary: the fully device-resident loader (clip bank in HBM as int16 PCM, batch gather and
    #      augmentation draws on the GPU, no H2D per step) - what a training loop would actually iterate ----
    bank = (torch.cat([h[0] for h in host]).clamp(-1, 1) * 32767).to(torch.int16).to(dev)     # RING*B clips
    dcfg = w.DrawConfig(seed=1, rir_prob=1.0, noise_prob=1.0)
    gidx = [torch.randperm(bank.shape[0], generator=torch.Generator().manual_seed(i))[:B].to(dev) for i in range(RING)]
    wv = torch.empty(B, N_SAMPLES, dtype=torch.float32, device=dev)

    def loader_step(i):
        w.gather_clips(bank, gidx[i % RING], out=wv)
        plan.featurize(wv, plan.draw_aug(dcfg, i * B, B, N_SAMPLES), out=out)

    for i in range(3):
        loader_step(i)
    barrier()
    e4, e5 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e4.record(stream)
    for i in range(args.steps):
        loader_step(i)
    e5.record(stream)
    barrier()
    loader_ms = max_over_ranks(e4.elapsed_time(e5))
    loader_value = world * B * args.steps / (loader_ms * 1e-3)

    # ---- supplementary: the waveform-shape augmentations of SURVEY.md section 8a row A3 (not part of configs[1]) ----
    g = torch.Generator().manual_seed(3)
    rates = (0.8 + 0.4 * torch.rand(B, generator=g, dtype=torch.float64)).to(dev)
    semis = torch.randint(-2, 3, (B,), generator=g, dtype=torch.int32).to(dev)
    acfg = w.DrawConfig(seed=1, rir_prob=1.0, noise_prob=1.0, stretch_prob=0.5, pitch_prob=0.5)
    shape_ops = {"time_stretch": lambda i: plan.time_stretch(dev_wav[i % RING], rates, rate_lo=0.8, out=wv),
                 "pitch_shift": lambda i: plan.pitch_shift(dev_wav[i % RING], semis, step_range=(-2, 2), out=wv),
                 "all_augmentations_pipeline": lambda i: plan.featurize(
                     dev_wav[i % RING], plan.draw_aug(acfg, i * B, B, N_SAMPLES), out=out)}
    shape_res = {}
    for name, fn in shape_ops.items():
        for i in range(3):
            fn(i)
        barrier()
        ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ea.record(stream)
        for i in range(args.steps):
            fn(i)
        eb.record(stream)
        barrier()
        ms = max_over_ranks(ea.elapsed_time(eb))
        shape_res[name] = {"value": world * B * args.steps / (ms * 1e-3), "unit": "clips/s", "ms_per_step": ms / args.steps}
    shape_res["what"] = ("supplementary: torchaudio-parity time-stretch (rate U[0.8,1.2) on every clip), pitch-shift (randint[-2,2] "
                         "semitones) and configs[1] with both drawn at probability 0.5 on the GPU in front of reverb + noise")

    # keep the GPU under the same load a little longer if the timed loops were too short to sample clocks
    if sampler.ok and len(sampler.samples) < 5:
        t_end = time.perf_counter() + 0.5
        i = 0
        while time.perf_counter() < t_end:
            step(i); i += 1
            if i % 64 == 0:
                torch.cuda.synchronize(dev)
        torch.cuda.synchronize(dev)
    clocks = sampler.stop()

    peak, peak_src = peaks()
    dom = max(kernel_ms, key=kernel_ms.get)          # the kernel with the longest average launch
    dom_ms = kernel_ms[dom]
    dom_bytes = KERNEL_BYTES[dom] * B
    achieved = dom_bytes / (dom_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": ncu_traffic(dom), "algorithmic_bytes": dom_bytes,
                "peak_source": peak_src,
                "kernel_ms": kernel_ms, "feature_stage_ms": feat_ms,
                "feature_path": "flat: feat_frames_kernel + feat_epilogue_mma_kernel (mix records from conv_kernel)" if n_split else "fused feat_kernel",
                "step_achieved": BYTES_STEP * B / (ms_step * 1e-3) / 1e9,
                "step_frac": BYTES_STEP * B / (ms_step * 1e-3) / 1e9 / peak,
                "bytes_per_clip": dict(KERNEL_BYTES, step=BYTES_STEP)}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cps, sec, cores = cpu_reference(256, reps=8, warm=1)
        cpu = {"value": cps, "unit": "clips/s", "cores": cores, "kind": "port",
               "sample": f"256 clips x 8 reps of the same workload ({sec * 1e3:.0f} ms each), oracle/ta_oracle.py "
                         "(torchaudio CPU, batched = the CPU's best case)",
               "per_clip_loop_value": cpu_per_clip_loop(), "per_clip_loop_sample": "24 clips, one call chain per clip "
               "(the reference's __getitem__ pattern)"}
        shape_res["cpu_baseline"] = cpu_shape_augs()

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": "clips/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": WORKLOAD, "batch_per_gpu": B, "global_batch": B * world, "n_samples": N_SAMPLES,
                           "noise_bank": f"{N_NOISE}x{NOISE_LEN}", "rir_bank": f"{N_RIR}x{RIR_LEN}",
                           "parallelism": f"clip-sharded x{world}, no collective",
                           "l2": f"ring of {RING} distinct input batches ({RING * B * N_SAMPLES * 4 / 1e6:.0f} MB) > 126 MB L2"},
                "e2e": {"value": e2e_value, "unit": "clips/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": e2e_ms / args.steps, "wall_ms_per_step": wall_ms / args.steps,
                        "numa_bound": numa_bound,
                        "pcie": {"h2d_copy_gbs": link_gbs, "e2e_h2d_gbs": e2e_h2d_gbs, "frac": e2e_h2d_gbs / link_gbs,
                                 "what": "bare pinned->device copy of one clip batch on this box vs the upload rate the e2e "
                                         "pipeline sustains: the host-fed path is bound by the PCIe link, not by the kernels"}},
                "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
                "e2e_pcm16": {"value": pcm_value, "unit": "clips/s", "ms_per_step": pcm_ms / args.steps,
                              "h2d_bytes_per_step": pinned_pcm[0].numel() * 2 + host_aug[0].nbytes(), "d2h_bytes_per_step": d2h,
                              "what": "supplementary: same as e2e but the host clips are int16 PCM (converted on the GPU, exact)"},
                "shape_augmentations": shape_res,
                "device_resident_loader": {"value": loader_value, "unit": "clips/s", "ms_per_step": loader_ms / args.steps,
                                           "what": "int16 PCM clip bank in HBM -> wwf_gather_clips -> wwf_draw_aug (on-GPU "
                                                   "Philox draws) -> wwf_featurize; no host->device copy per step"}}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
