#!/usr/bin/env python
"""bench.py - headline benchmark of the B200 audio feature path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config cfg1|cfg2|cfg3|cfg4|cfg5]

Metric (BASELINE.json): featurized clips/s, augmentation + features.  The default workload (cfg2) is BASELINE.json
configs[1] - the configuration the metric is quoted on for one GPU: MFCC-40 (n_fft 400, hop 160, 40 mels) +
background-noise mixing at a target SNR + RIR reverb, batch 1024 synthetic 1.5 s 16 kHz clips PER GPU (weak scaling;
the batch is sharded by clip index, no collective on the feature path).  One "step" = one pass of the hot path over
one batch.  Prints ONE JSON line (rank 0).

  value        whole-job clips/s with the clips already resident in HBM (CUDA events, max over ranks)
  e2e          same metric through the public API with HOST buffers: pinned host clips + draws
               -> H2D -> wwf_featurize -> D2H of the features, all inside the timed region
  roofline     dominant kernel: algorithmic bytes / its CUDA-event launch time vs measured HBM peak
  parity       BEFORE timing: 32 clips sampled from the first batch's output (the bench's own launch shape) against the
               oracle on the same inputs; the run fails above tolerance
  cpu_baseline the oracle (torchaudio CPU, the reference's arithmetic) on a bounded sample, rank 0, N=1

--config selects another BASELINE.json configuration (cfgK = configs[K-1]):
  cfg1  40-bin log-mel (n_fft 400, hop 160), batch 64 x 1.5 s, no augmentation (the reference's CPU-runnable case)
  cfg3  Default preset (augment + log-mel-128 + SpecAugment) feeding the reference's ResNet-18 train step under DDP:
        metric = train samples/s, with the feature stage's share of the step and the gradient all-reduce cost
  cfg4  Edge deployment: 2 s clips, 64 mels, float16 features, noise 0.5 / RIR 0.3
  cfg5  Large-dataset sweep: 1 M synthetic 2 s clips streamed from a device ring at batch 256 ... 8192
--impl reference times the torchaudio CPU path alone, on all host threads, on the same workload and batch.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import sys
import threading
import time
from dataclasses import dataclass

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

SR = 16000


# ---- workloads: BASELINE.json configs -------------------------------------------------------------------------
@dataclass
class Workload:
    key: str
    title: str
    feature_type: str
    n_fft: int
    hop: int
    n_mels: int
    n_mfcc: int
    n_samples: int
    batch: int                    # clips per GPU per step
    f16: bool = False
    rir_prob: float = 0.0         # 1.0 = every clip (configs[1]); < 1: Bernoulli per clip, -1 = dry
    noise_prob: float = 0.0
    mask_prob: float = 0.0        # SpecAugment: 2 frequency masks (param 15) + 2 time masks (param 35) per selected clip
    n_noise: int = 256
    n_rir: int = 64
    rir_len: int = 8000

    @property
    def T(self):
        return self.n_samples // self.hop + 1

    @property
    def F(self):
        return self.n_mfcc if self.feature_type == "mfcc" else self.n_mels

    @property
    def out_bytes(self):
        return (2 if self.f16 else 4) * self.F * self.T

    @property
    def ring(self):
        """distinct input batches cycled through: together larger than the 126 MB L2"""
        return max(4, math.ceil(160e6 / (self.batch * self.n_samples * 4)))

    def bytes_per_clip(self):
        """algorithmic (compulsory) bytes per clip, SURVEY.md section 8d: every input sample read once, every output
        element written once; expected values over the augmentation probabilities."""
        n4, m = 4 * self.n_samples, 4 * self.n_mels * self.T
        step = n4 + self.noise_prob * n4 + self.rir_prob * 4 * self.rir_len + self.out_bytes
        return {"step": step,
                "conv_kernel": self.rir_prob * (n4 + 4 * self.rir_len + n4),                  # x, h in; y out
                "feat_kernel": n4 + self.noise_prob * n4 + self.out_bytes,                    # y, noise in; features out
                "feat_frames_kernel": n4 + self.noise_prob * n4 + m,                          # y, noise in; dB tile out
                "feat_epilogue_block_kernel": m + self.out_bytes, "feat_epilogue_mma_kernel": m + self.out_bytes,
                "feat_prep_kernel": 0}


WORKLOADS = {
    "cfg1": Workload("cfg1", "configs[0]: 40-bin log-mel (n_fft 400, hop 160), batch 64 x 1.5 s @ 16 kHz per GPU, no augmentation",
                     "mel", 400, 160, 40, 40, 24000, 64),
    "cfg2": Workload("cfg2", "configs[1]: MFCC-40 (n_fft 400, hop 160, 40 mels) + noise@SNR U[5,20] dB + RIR reverb "
                     "(8000 taps), batch 1024 x 1.5 s @ 16 kHz per GPU", "mfcc", 400, 160, 40, 40, 24000, 1024,
                     rir_prob=1.0, noise_prob=1.0),
}
WORKLOADS["cfg4"] = Workload("cfg4", "configs[3]: Edge deployment front end - log-mel-64 (n_fft 1024, hop 160), float16 features, "
                             "2 s clips, noise 0.5 @ SNR U[5,20] dB, RIR 0.3, SpecAugment 0.5, batch 1024 per GPU",
                             "mel", 1024, 160, 64, 32, 32000, 1024, f16=True, rir_prob=0.3, noise_prob=0.5, mask_prob=0.5)
WORKLOADS["cfg5"] = Workload("cfg5", "configs[4]: Large-dataset sweep - 1 M synthetic 2 s clips streamed from a device ring through "
                             "log-mel-40 (n_fft 400, hop 160), batch 256...8192 (headline batch 4096) per GPU", "mel", 400, 160, 40, 40,
                             32000, 4096)
CFG3_TITLE = ("configs[2]: Default preset (noise 0.5 @ SNR U[5,20] dB, RIR 0.25, SpecAugment 0.5; log-mel-128, n_fft 1024, hop 160, "
              "1.5 s clips) -> reference ResNet-18 train step (AdamW, fp32), DistributedDataParallel, batch 128 per GPU")
METRIC = "featurized clips/sec (1.5s@16kHz, aug+log-mel+DCT: configs[1] MFCC-40 + noise@SNR + RIR)"
# the default workload's numbers under their old names (tools/*.py)
_D = WORKLOADS["cfg2"]
N_FFT, HOP, N_MELS, N_MFCC, B_PER_GPU, N_SAMPLES = _D.n_fft, _D.hop, _D.n_mels, _D.n_mfcc, _D.batch, _D.n_samples


def metric_name(wl: Workload) -> str:
    if wl.key == "cfg2":
        return METRIC
    return f"featurized clips/sec ({wl.n_samples / SR:g}s@16kHz, {wl.key})"


def config_dict(wl: Workload, world: int) -> dict:
    """The `config` object of the JSON line - identical for our arm and the reference arm."""
    ring = wl.ring
    return {"workload": wl.title, "batch_per_gpu": wl.batch, "global_batch": wl.batch * world, "n_samples": wl.n_samples,
            "noise_bank": f"{wl.n_noise}x{wl.n_samples}", "rir_bank": f"{wl.n_rir}x{wl.rir_len}",
            "parallelism": f"clip-sharded x{world}, no collective",
            "l2": f"ring of {ring} distinct input batches ({ring * wl.batch * wl.n_samples * 4 / 1e6:.0f} MB) > 126 MB L2"}


def synth(seed: int, B: int, wl: Workload = None):
    """One synthetic batch and its explicit augmentation draws (host tensors); default workload: configs[1]."""
    wl = wl or WORKLOADS["cfg2"]
    g = torch.Generator().manual_seed(seed)
    wav = 0.1 * torch.randn(B, wl.n_samples, generator=g)
    draws = {}
    if wl.rir_prob > 0 or wl.noise_prob > 0:
        draws = dict(rir_idx=torch.randint(0, wl.n_rir, (B,), generator=g, dtype=torch.int32),
                     noise_idx=torch.randint(0, wl.n_noise, (B,), generator=g, dtype=torch.int32),
                     noise_off=torch.randint(0, wl.n_samples, (B,), generator=g),
                     snr_db=5.0 + 15.0 * torch.rand(B, generator=g))
        if wl.rir_prob < 1.0:
            draws["rir_idx"][torch.rand(B, generator=g) >= wl.rir_prob] = -1
        if wl.noise_prob < 1.0:
            draws["noise_idx"][torch.rand(B, generator=g) >= wl.noise_prob] = -1
    if wl.mask_prob > 0:
        import wakeword_trainer_home_b200.pipeline as P
        fs, fl = P.draw_mask_params(g, B, wl.F, 15, 2)
        ts, tl = P.draw_mask_params(g, B, wl.T, 35, 2)
        off = torch.rand(B, generator=g) >= wl.mask_prob
        fl[off] = 0
        tl[off] = 0
        draws.update(fmask_start=fs, fmask_len=fl, tmask_start=ts, tmask_len=tl)
    return wav, draws


def synth_banks(wl: Workload = None):
    wl = wl or WORKLOADS["cfg2"]
    g = torch.Generator().manual_seed(1234)
    noise = [0.05 * torch.randn(wl.n_samples, generator=g) for _ in range(wl.n_noise)]
    t = torch.arange(wl.rir_len, dtype=torch.float32)
    rirs = [torch.randn(wl.rir_len, generator=g) * torch.exp(-t / 1000.0) for _ in range(wl.n_rir)]
    return noise, rirs


def make_plan(w, wl: Workload, dev):
    plan = w.FeaturePlan(SR, wl.feature_type, wl.n_mels, wl.n_mfcc, wl.n_fft, wl.hop, dev,
                         out_dtype=torch.float16 if wl.f16 else torch.float32,
                         n_freq_masks=2 if wl.mask_prob > 0 else 0, n_time_masks=2 if wl.mask_prob > 0 else 0)
    banks = (None, None)
    if wl.rir_prob > 0 or wl.noise_prob > 0:
        banks = synth_banks(wl)
        plan.register_noise(banks[0])
        plan.register_rirs(banks[1])
    return plan, banks


def oracle_kwargs(wl: Workload, d: dict, banks, sel=None):
    pick = (lambda v: v) if sel is None else (lambda v: v[sel])
    kw = dict(sample_rate=SR, feature_type=wl.feature_type, n_mels=wl.n_mels, n_mfcc=wl.n_mfcc, n_fft=wl.n_fft, hop_length=wl.hop)
    if "rir_idx" in d:
        kw.update(rirs=banks[1], rir_idx=pick(d["rir_idx"]), noise_bank=banks[0], noise_idx=pick(d["noise_idx"]),
                  noise_off=pick(d["noise_off"]), snr_db=pick(d["snr_db"]))
    if "fmask_start" in d:
        kw.update(fstart=pick(d["fmask_start"]), flen=pick(d["fmask_len"]), tstart=pick(d["tmask_start"]), tlen=pick(d["tmask_len"]))
    return kw


def ncu_traffic(kernel: str):
    """DRAM bytes per launch of `kernel` on the cfg2 workload from the committed ncu capture (newest profiles/rNN_traffic.json)."""
    for name in ("r02_traffic.json", "r01_traffic.json"):
        try:
            with open(os.path.join(ROOT, "profiles", name)) as f:
                return int(json.load(f)[kernel])
        except Exception:
            continue
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
def cpu_reference(wl: Workload, n_clips: int, reps: int, warm: int):
    """Times the oracle (torchaudio CPU) on n_clips clips of the workload; returns (clips/s, s per rep, threads)."""
    from oracle import ta_oracle as tao
    torch.set_num_threads(os.cpu_count() or 1)
    banks = synth_banks(wl) if (wl.rir_prob > 0 or wl.noise_prob > 0) else (None, None)
    wav, d = synth(7, n_clips, wl)
    kw = oracle_kwargs(wl, d, banks)
    for _ in range(warm):
        tao.pipeline(wav, **kw)
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        tao.pipeline(wav, **kw)
        ts.append(time.perf_counter() - t0)
    return n_clips * len(ts) / sum(ts), sum(ts) / len(ts), torch.get_num_threads()


def cpu_per_clip_loop(wl: Workload, n_clips: int = 24):
    """The reference's real usage pattern (one call chain per clip, src/evaluation/evaluator.py:202-205,
    Dataset.__getitem__): clips/s of the oracle called clip by clip."""
    from oracle import ta_oracle as tao
    banks = synth_banks(wl) if (wl.rir_prob > 0 or wl.noise_prob > 0) else (None, None)
    wav, d = synth(8, n_clips, wl)

    def one(i):
        tao.pipeline(wav[i:i + 1], **oracle_kwargs(wl, d, banks, slice(i, i + 1)))
    one(0)
    t0 = time.perf_counter()
    for i in range(n_clips):
        one(i)
    return n_clips / (time.perf_counter() - t0)


def cpu_shape_augs(n_samples: int, n_stretch: int = 32, n_pitch: int = 8):
    """torchaudio CPU (oracle/ta_oracle.py) on a bounded sample: time-stretch and F.pitch_shift, clips/s."""
    from oracle import ta_oracle as tao
    g = torch.Generator().manual_seed(3)
    x = 0.1 * torch.randn(max(n_stretch, n_pitch), n_samples, generator=g)
    rates = 0.8 + 0.4 * torch.rand(n_stretch, generator=g, dtype=torch.float64)
    t0 = time.perf_counter()
    tao.time_stretch(x[:n_stretch], rates)
    t1 = time.perf_counter()
    tao.pitch_shift(x[:n_pitch], torch.tensor([-2, -1, 1, 2] * (n_pitch // 4), dtype=torch.int32), 16000)
    t2 = time.perf_counter()
    return {"time_stretch": n_stretch / (t1 - t0), "pitch_shift": n_pitch / (t2 - t1), "unit": "clips/s",
            "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{n_stretch} clips time-stretch, {n_pitch} clips F.pitch_shift (1.5 s each), torchaudio CPU"}


def cfg3_workload() -> Workload:
    return Workload("cfg3", CFG3_TITLE, "mel", 1024, 160, 128, 40, 24000, 128, rir_prob=0.25, noise_prob=0.5, mask_prob=0.5,
                    n_noise=64, n_rir=16)


def run_reference(args, rank: int, world: int):
    """The reference's torchaudio CPU arithmetic on the same workload, batch and config keys as our arm (rank 0 only)."""
    if rank != 0:
        return
    wl = cfg3_workload() if args.config == "cfg3" else WORKLOADS[args.config]
    n_clips = wl.batch
    cps, sec, cores = cpu_reference(wl, n_clips, reps=args.steps, warm=max(1, min(args.warmup, 3)))
    cfg = config_dict(wl, world)
    sample = (f"{n_clips} clips/step x {args.steps} steps; oracle/ta_oracle.py = the reference's torchaudio CPU arithmetic "
              "(its src/data module is absent upstream), one batched call chain per step (fftconvolve per RIR group, add_noise, "
              "MelSpectrogram/MFCC) - the CPU's best case")
    if args.config == "cfg3":
        sample += "; feature side of the train step only (what the reference's DataLoader workers compute)"
    line = {"impl": "reference", "metric": metric_name(wl) if args.config != "cfg3" else "train samples/sec (feature side, CPU)",
            "value": cps, "unit": "clips/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": cfg,
            "cpu_baseline": {"value": cps, "unit": "clips/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": cps, "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---- parity gate ---------------------------------------------------------------------------------
ABS_DB, REL = 1e-3, 1e-4          # north_star: max-abs 1e-3 dB, relative 1e-4 (tests/helpers.py)


def parity_gate(wl: Workload, out: torch.Tensor, wav: torch.Tensor, draws: dict, banks, n: int = 32):
    """Compare n clips sampled with a stride from a featurized batch (the bench's own launch shape) with the oracle on
    the same inputs and draws.  float16 output is held to half an ulp of half on top of the float32 bound."""
    from oracle import ta_oracle as tao
    B = wav.shape[0]
    sel = torch.arange(0, B, max(1, B // n))[:n]
    ref = tao.pipeline(wav[sel], **oracle_kwargs(wl, draws, banks, sel)).double()
    got = out[sel.to(out.device)].cpu().double()
    err = (got - ref).abs()
    bound = ABS_DB + REL * ref.abs()
    if wl.f16:
        bound = bound + ref.abs() * 2.0 ** -11 + 2.0 ** -24
    rel_l2 = float((got - ref).norm() / ref.norm().clamp_min(1e-30))
    ok = bool((err <= bound).all()) and bool(torch.isfinite(got).all()) and rel_l2 <= (6e-4 if wl.f16 else REL)
    masked = int((ref == 0).sum()) if wl.mask_prob > 0 else 0
    return {"ok": ok, "n": int(sel.numel()), "max_abs_db": float(err.max()), "rel_l2": rel_l2,
            "tolerance": f"|d| <= {ABS_DB} + {REL}|ref|" + (" + half ulp(f16)" if wl.f16 else "") + ", rel_l2 <= " + ("6e-4" if wl.f16 else str(REL)),
            "masked_elements_equal": bool((got[ref == 0] == 0).all()) if masked else None,
            "against": "oracle/ta_oracle.py (torchaudio CPU) on the same clips and draws"}


# ---- our arm: feature workloads -------------------------------------------------------------------------------
def run_features(args, wl: Workload, rank: int, local_rank: int, world: int):
    import torch.distributed as dist
    import wakeword_trainer_home_b200 as w

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

    default = wl.key == "cfg2"
    B, N, RING = wl.batch, wl.n_samples, wl.ring
    plan, banks = make_plan(w, wl, dev)
    host = [synth(100 * rank + i, B, wl) for i in range(RING)]
    has_aug = bool(host[0][1])
    odt = torch.float16 if wl.f16 else torch.float32
    # pinned host buffers are allocated (and first touched) on the cores next to this GPU's PCIe root; a
    # single-GPU run gets its full CPU affinity back afterwards so the CPU baseline still uses every core
    with numa_local(local_rank) as nl:
        pinned_wav = [h[0].pin_memory() for h in host]
        pinned_draws = [{k: v.pin_memory() for k, v in h[1].items()} for h in host]
        sf = w.StreamedFeaturizer(plan, B, N, depth=2, copy_back=True)
        if default:
            pinned_pcm = [(h.clamp(-1, 1) * 32767).to(torch.int16).pin_memory() for h in pinned_wav]
            sf16 = w.StreamedFeaturizer(plan, B, N, depth=2, copy_back=True, pcm16=True)
    numa_bound = numa_bound or nl.bound
    dev_wav = [h.to(dev) for h in pinned_wav]
    dev_aug = [w.AugParams(**d).to(dev) if has_aug else None for d in pinned_draws]
    out = torch.empty(B, 1, wl.F, wl.T, dtype=odt, device=dev)
    stream = torch.cuda.current_stream(dev)

    def step(i):
        plan.featurize(dev_wav[i % RING], dev_aug[i % RING], out=out)

    # ---- parity gate: the first batch, featurized with the launch shape the timed loop uses, vs the oracle ----
    step(0)
    torch.cuda.synchronize(dev)
    parity = parity_gate(wl, out, host[0][0], host[0][1], banks) if rank == 0 else None
    if parity is not None and not parity["ok"]:
        print(json.dumps({"error": "parity gate failed before timing", "parity": parity}), flush=True)
        raise SystemExit(3)

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
    feat_ms = sum(v for k, v in kernel_ms.items() if k != "conv_kernel")

    # ---- end to end through the public API with HOST buffers (pinned), copies inside the timed
    #      region: upload of clips + draws, featurize, download of the features, triple-streamed ----
    host_aug = [w.AugParams(**d) if has_aug else None for d in pinned_draws]
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
    h2d = pinned_wav[0].numel() * 4 + (host_aug[0].nbytes() if has_aug else 0)
    d2h = sf.h_out[0].numel() * sf.h_out[0].element_size()
    # what the link gives: a bare pinned -> device copy of one clip batch, best of 5 (alone on the box: only this rank copies)
    link_ms = 1e9
    for r in range(world):
        barrier()
        if r == rank:
            for _ in range(5):
                ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                ea.record(stream)
                dev_wav[0].copy_(pinned_wav[0], non_blocking=True)
                eb.record(stream)
                eb.synchronize()
                link_ms = min(link_ms, ea.elapsed_time(eb))
        if world > 8:
            break
    link_gbs = pinned_wav[0].numel() * 4 / (link_ms * 1e-3) / 1e9
    # ... and what it gives when EVERY rank copies at once, in both directions like the e2e pipeline does (clips up on
    # one stream, features down on another): the ceiling of the host-fed path on a box whose GPUs share host DRAM / PCIe roots
    reps = 8
    barrier()
    ca, cb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    cc, cd = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ca.record(sf.s_in)
    cc.record(sf.s_out)
    for _ in range(reps):
        with torch.cuda.stream(sf.s_in):
            sf.d_wav[0].copy_(pinned_wav[0], non_blocking=True)
        with torch.cuda.stream(sf.s_out):
            sf.h_out[0].copy_(sf.d_out[0], non_blocking=True)
    cb.record(sf.s_in)
    cd.record(sf.s_out)
    sf.synchronize()
    conc_h2d_ms = max_over_ranks(ca.elapsed_time(cb)) / reps
    conc_d2h_ms = max_over_ranks(cc.elapsed_time(cd)) / reps
    conc_h2d_gbs = pinned_wav[0].numel() * 4 / (conc_h2d_ms * 1e-3) / 1e9
    conc_d2h_gbs = d2h / (conc_d2h_ms * 1e-3) / 1e9
    e2e_h2d_gbs = h2d / (e2e_ms / args.steps * 1e-3) / 1e9
    barrier()

    extra = {}
    if default:
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
        extra["e2e_pcm16"] = {"value": world * B * args.steps / (pcm_ms * 1e-3), "unit": "clips/s", "ms_per_step": pcm_ms / args.steps,
                              "h2d_bytes_per_step": pinned_pcm[0].numel() * 2 + host_aug[0].nbytes(), "d2h_bytes_per_step": d2h,
                              "what": "supplementary: same as e2e but the host clips are int16 PCM (the native format of WAV files; "
                                      "converted on the GPU, exact) - the recommended host format of StreamedFeaturizer / GpuBatchLoader"}

        # ---- supplementary: the fully device-resident loader (clip bank in HBM as int16 PCM, batch gather and
        #      augmentation draws on the GPU, no H2D per step) - what a training loop would actually iterate ----
        bank = (torch.cat([h[0] for h in host]).clamp(-1, 1) * 32767).to(torch.int16).to(dev)     # RING*B clips
        dcfg = w.DrawConfig(seed=1, rir_prob=1.0, noise_prob=1.0)
        gidx = [torch.randperm(bank.shape[0], generator=torch.Generator().manual_seed(i))[:B].to(dev) for i in range(RING)]
        wv = torch.empty(B, N, dtype=torch.float32, device=dev)

        def loader_step(i):
            w.gather_clips(bank, gidx[i % RING], out=wv)
            plan.featurize(wv, plan.draw_aug(dcfg, i * B, B, N), out=out)

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
        extra["device_resident_loader"] = {"value": world * B * args.steps / (loader_ms * 1e-3), "unit": "clips/s",
                                           "ms_per_step": loader_ms / args.steps,
                                           "what": "int16 PCM clip bank in HBM -> wwf_gather_clips -> wwf_draw_aug (on-GPU "
                                                   "Philox draws) -> wwf_featurize; no host->device copy per step"}

        # ---- supplementary: the waveform-shape augmentations of SURVEY.md section 8a row A3 (not part of configs[1]) ----
        g = torch.Generator().manual_seed(3)
        rates = (0.8 + 0.4 * torch.rand(B, generator=g, dtype=torch.float64)).to(dev)
        semis = torch.randint(-2, 3, (B,), generator=g, dtype=torch.int32).to(dev)
        acfg = w.DrawConfig(seed=1, rir_prob=1.0, noise_prob=1.0, stretch_prob=0.5, pitch_prob=0.5)
        shape_ops = {"time_stretch": lambda i: plan.time_stretch(dev_wav[i % RING], rates, rate_lo=0.8, out=wv),
                     "pitch_shift": lambda i: plan.pitch_shift(dev_wav[i % RING], semis, step_range=(-2, 2), out=wv),
                     "all_augmentations_pipeline": lambda i: plan.featurize(
                         dev_wav[i % RING], plan.draw_aug(acfg, i * B, B, N), out=out)}
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
        extra["shape_augmentations"] = shape_res

    if wl.key == "cfg5":
        # ---- the sweep itself: 1 M clips streamed from a device ring of 65 536 clips (8.4 GB) at every batch size ----
        ring_clips = 65536
        bank = torch.empty(ring_clips, N, dtype=torch.float32, device=dev)
        for a in range(0, ring_clips, 8192):
            bank[a:a + 8192].normal_(0.0, 0.1)
        total = 1_000_000
        sweep = {}
        for bs in (256, 512, 1024, 2048, 4096, 8192):
            o = torch.empty(bs, 1, wl.F, wl.T, dtype=odt, device=dev)
            nsteps = (total + bs - 1) // bs
            slots = ring_clips // bs
            for i in range(3):
                plan.featurize(bank[i * bs:(i + 1) * bs], None, out=o)
            barrier()
            ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ea.record(stream)
            for i in range(nsteps):
                s = (i % slots) * bs
                plan.featurize(bank[s:s + bs], None, out=o)
            eb.record(stream)
            barrier()
            ms = max_over_ranks(ea.elapsed_time(eb))
            sweep[str(bs)] = {"clips_per_s": world * bs * nsteps / (ms * 1e-3), "ms_per_step": ms / nsteps, "steps": nsteps,
                              "hbm_gbs": world * bs * nsteps * wl.bytes_per_clip()["step"] / (ms * 1e-3) / 1e9}
            del o
        extra["sweep"] = {"clips_streamed_per_batch_size": total, "ring": f"{ring_clips} clips x {N} samples f32 in HBM "
                          f"({ring_clips * N * 4 / 1e9:.1f} GB), cycled", "by_batch": sweep}
        del bank

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
    kb = wl.bytes_per_clip()
    dom = max(kernel_ms, key=kernel_ms.get)          # the kernel with the longest average launch
    dom_ms = kernel_ms[dom]
    dom_bytes = kb[dom] * B
    achieved = dom_bytes / (dom_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": ncu_traffic(dom) if default else None, "algorithmic_bytes": dom_bytes,
                "peak_source": peak_src,
                "kernel_ms": kernel_ms, "feature_stage_ms": feat_ms,
                "feature_path": "flat: feat_frames_kernel + epilogue kernel" if n_split else "fused feat_kernel",
                "step_achieved": kb["step"] * B / (ms_step * 1e-3) / 1e9,
                "step_frac": kb["step"] * B / (ms_step * 1e-3) / 1e9 / peak,
                "bytes_per_clip": kb}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        n_cpu = min(256, B)
        cps, sec, cores = cpu_reference(wl, n_cpu, reps=8, warm=1)
        cpu = {"value": cps, "unit": "clips/s", "cores": cores, "kind": "port",
               "sample": f"{n_cpu} clips x 8 reps of the same workload ({sec * 1e3:.0f} ms each), oracle/ta_oracle.py "
                         "(torchaudio CPU, batched = the CPU's best case)",
               "per_clip_loop_value": cpu_per_clip_loop(wl), "per_clip_loop_sample": "24 clips, one call chain per clip "
               "(the reference's __getitem__ pattern)"}
        if default:
            extra["shape_augmentations"]["cpu_baseline"] = cpu_shape_augs(N)

    if rank == 0:
        line = {"metric": metric_name(wl), "value": value, "unit": "clips/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": config_dict(wl, world),
                "e2e": {"value": e2e_value, "unit": "clips/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": e2e_ms / args.steps, "wall_ms_per_step": wall_ms / args.steps,
                        "numa_bound": numa_bound,
                        "pcie": {"h2d_copy_gbs": link_gbs, "e2e_h2d_gbs": e2e_h2d_gbs, "frac": e2e_h2d_gbs / link_gbs,
                                 "concurrent_h2d_gbs": conc_h2d_gbs, "concurrent_d2h_gbs": conc_d2h_gbs,
                                 "frac_of_concurrent": e2e_h2d_gbs / conc_h2d_gbs,
                                 "what": "per GPU: bare pinned->device copy of one clip batch with only one rank copying (h2d_copy_gbs), "
                                         "the same copy with EVERY rank uploading clips and downloading features at once "
                                         "(concurrent_*, max over ranks) and the upload rate the e2e pipeline sustains: the host-fed "
                                         "path is bound by the box's PCIe / host-memory system, not by the kernels"}},
                "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "parity": parity, "cpu_baseline": cpu}
        line.update(extra)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ---- our arm: configs[2], the feature path feeding the reference's train step under DDP -----------------------
def run_cfg3(args, rank: int, local_rank: int, world: int):
    import torch.distributed as dist
    import wakeword_trainer_home_b200 as w
    from wakeword_trainer_home_b200 import ddp_training as dt
    from wakeword_trainer_home_b200.loader import GpuBatchLoader

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29533")
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    import logging
    logging.disable(logging.WARNING)

    def barrier():
        dist.barrier(device_ids=[local_rank])
        torch.cuda.synchronize(dev)

    def max_over_ranks(ms: float) -> float:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    wl = cfg3_workload()
    cfg = dt.default_config()                      # WakewordConfig(): the Default preset's values
    cfg.data.audio_duration = wl.n_samples / SR    # BASELINE's 1.5 s clips (the preset's own 2.5 s: --clip-seconds 2.5)
    if args.clip_seconds:
        cfg.data.audio_duration = args.clip_seconds
    N = int(SR * cfg.data.audio_duration)
    B = int(cfg.training.batch_size)               # 128 per GPU (src/config/defaults.py:35)
    K, W = args.steps, args.warmup
    noise, rirs = dt.synthetic_aug_banks(wl.n_noise, N, wl.n_rir, wl.rir_len)
    plan, train, _ = dt.build_plan_and_loaders(cfg, dev, rank=rank, world_size=world, n_train_clips=(K + W) * B * world,
                                               n_samples=N, seed=0, noise=noise, rirs=rirs)
    T, F = plan.num_frames(N), plan.n_feat
    timed = dt.TimedLoader(train)
    import contextlib
    with contextlib.redirect_stdout(sys.stderr):      # the reference prints a CUDA banner: keep stdout to the one JSON line
        trainer = dt.build_ddp_trainer(cfg, timed, [], dev, local_rank=local_rank)
    gbytes = dt.grad_bytes(trainer.model)

    class Steps:
        """K batches of an underlying loader per 'epoch' (Trainer.train_epoch runs the whole iterable)."""
        def __init__(self, it, k):
            self.it, self.k = it, k
        def __len__(self):
            return self.k
        def __iter__(self):
            for _ in range(self.k):
                yield next(self.it)

    n0 = w.launch_count()
    it = iter(timed)
    trainer.train_loader = Steps(it, W)
    trainer.train_epoch(0)                         # warm-up: cuDNN autotune, DDP bucket rebuild, allocator
    timed.feature_ms()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    trainer.train_loader = Steps(it, K)
    l0 = w.launch_count()
    e0.record()
    t0 = time.perf_counter()
    loss, acc = trainer.train_epoch(1)             # exactly K optimizer steps of the reference's loop
    e1.record()
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    launches = w.launch_count() - l0
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    feat_ms = max_over_ranks(timed.feature_ms())
    value = world * B * K / (ms_total * 1e-3)
    ar_ms = dt.time_allreduce(gbytes, dev) if world > 1 else 0.0

    # feature stage alone on the same loader shape (no model): what the path sustains when the classifier is not the limit
    solo = iter(dt.build_plan_and_loaders(cfg, dev, rank=rank, world_size=world, n_train_clips=(K + 3) * B * world, n_samples=N,
                                          seed=1, noise=noise, rirs=rirs)[1])
    for _ in range(3):
        next(solo)
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    for _ in range(K):
        next(solo)
    e3.record()
    barrier()
    solo_ms = max_over_ranks(e2.elapsed_time(e3))

    # e2e: the same train step fed from PINNED HOST clips (H2D of every batch inside the timed region, loss read back per step)
    g = torch.Generator().manual_seed(5)
    hclips = (0.1 * torch.randn((K + 3) * B, N, generator=g)).pin_memory()
    hlabels = torch.randint(0, 2, ((K + 3) * B,), generator=g)
    hl = GpuBatchLoader(hclips, hlabels, plan, B, augment=None, spec_augment=None, shuffle=False, rank=0, world_size=1)
    hit = iter(hl)
    trainer.train_loader = Steps(hit, 3)
    trainer.train_epoch(2)
    barrier()
    e4, e5 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    trainer.train_loader = Steps(hit, K)
    e4.record()
    trainer.train_epoch(3)
    e5.record()
    barrier()
    e2e_ms = max_over_ranks(e4.elapsed_time(e5))

    if rank == 0:
        bytes_clip = wl.bytes_per_clip()["step"] if N == wl.n_samples else None
        line = {"metric": "train samples/sec (Default preset features on the B200 path -> reference ResNet-18 train step, DDP)",
                "value": value, "unit": "samples/s", "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_total / K,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": CFG3_TITLE, "batch_per_gpu": B, "global_batch": B * world, "n_samples": N,
                           "features": f"(B,1,{F},{T}) float32", "model": cfg.model.architecture, "optimizer": cfg.optimizer.optimizer,
                           "mixed_precision": bool(cfg.optimizer.mixed_precision),
                           "parallelism": f"ddp{world}: clip-sharded feature path (no collective) + NCCL gradient all-reduce",
                           "loader": "DeviceBatchLoader: int16 PCM clip bank in HBM, on-GPU gather + Philox draws + wwf_featurize",
                           "trainer": "reference src/training/trainer.py Trainer.train_epoch, unmodified"},
                "train": {"loss": loss, "accuracy": acc, "finite": bool(math.isfinite(loss)), "wall_ms_per_step": wall_ms / K},
                "feature_stage": {"ms_per_step": feat_ms / K, "share_of_step": feat_ms / ms_total,
                                  "alone_ms_per_step": solo_ms / K, "alone_clips_per_s": world * B * K / (solo_ms * 1e-3),
                                  "bytes_per_clip": bytes_clip,
                                  "what": "CUDA events around every batch the loader produced inside the timed train loop; "
                                          "'alone' = the same loader iterated without the classifier"},
                "collective": {"kind": "NCCL all-reduce of the gradients (DistributedDataParallel)", "bytes_per_step": gbytes,
                               "allreduce_ms": ar_ms, "busbw_gbs": (2 * (world - 1) / world * gbytes / (ar_ms * 1e-3) / 1e9) if ar_ms else None,
                               "share_of_step_if_exposed": ar_ms / (ms_total / K) if ar_ms else 0.0},
                "e2e": {"value": world * B * K / (e2e_ms * 1e-3), "unit": "samples/s", "h2d_bytes_per_step": B * N * 4 + B * 8,
                        "d2h_bytes_per_step": 8, "ms_per_step": e2e_ms / K,
                        "what": "same train step with every batch uploaded from pinned host clips inside the timed region "
                                "(GpuBatchLoader) and the loss read back per step (Trainer's loss.item())"},
                "gpu_launches": int(launches), "cpu_baseline": None, "roofline": None}
        print(json.dumps(line), flush=True)
    dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="cfg2", choices=["cfg1", "cfg2", "cfg3", "cfg4", "cfg5"])
    ap.add_argument("--clip-seconds", type=float, default=0.0, help="cfg3 only: clip length (the Default preset's own is 2.5)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the feature path has no CPU fallback")
    if args.config == "cfg3":
        run_cfg3(args, rank, local_rank, world)
    else:
        run_features(args, WORKLOADS[args.config], rank, local_rank, world)


if __name__ == "__main__":
    main()
