"""``AudioAugmentation`` and ``SpecAugment`` - drop-ins for the reference's
``src.data.augmentation`` classes (module absent from the reference checkout; surface from
tests/test_training_pipeline.py:230-262 and the kwargs dict at src/ui/panel_training.py:309-318).

Both classes only *draw* parameters on the host (torch.Generator) and hand them, explicit,
to the CUDA path; ``draw()`` is public so a test can feed the very same draws to the oracle.

Time-stretch and pitch-shift (``time_stretch_range``, ``pitch_shift_range``)
are drawn per clip with probabilities ``time_stretch_prob`` / ``pitch_shift_prob`` - the reference's module
is absent, so its apply probabilities are unknown; 0.5 is this package's choice - and applied before the
reverb: time-stretch -> pitch-shift -> RIR -> noise.
"""
from __future__ import annotations

from typing import Optional, Sequence

import torch

from .pipeline import AugParams, FeaturePlan, draw_mask_params, spec_augment_


class AudioAugmentation:
    def __init__(self, sample_rate: int = 16000, device: str = "cuda", time_stretch_range=(0.8, 1.2),
                 pitch_shift_range=(-2, 2), background_noise_prob: float = 0.5, noise_snr_range=(5.0, 20.0),
                 rir_prob: float = 0.25, background_noise: Optional[Sequence[torch.Tensor]] = None,
                 rirs: Optional[Sequence[torch.Tensor]] = None, seed: Optional[int] = None,
                 plan: Optional[FeaturePlan] = None, time_stretch_prob: float = 0.5, pitch_shift_prob: float = 0.5):
        self.sample_rate = sample_rate
        self.time_stretch_range = (float(time_stretch_range[0]), float(time_stretch_range[1]))
        self.pitch_shift_range = (int(pitch_shift_range[0]), int(pitch_shift_range[1]))   # integer semitones, validator.py:289-294
        self.time_stretch_prob, self.pitch_shift_prob = float(time_stretch_prob), float(pitch_shift_prob)
        self.background_noise_prob = float(background_noise_prob)
        self.noise_snr_range = tuple(noise_snr_range)
        self.rir_prob = float(rir_prob)
        # the time-domain kernels hang off a plan; feature settings are irrelevant for augment()
        self.plan = plan if plan is not None else FeaturePlan(sample_rate=sample_rate, n_fft=400, hop_length=160,
                                                              n_mels=40, device=device)
        self.device = self.plan.device
        self.noise_lens = [int(c.numel()) for c in background_noise] if background_noise else []
        if background_noise:
            self.plan.register_noise(background_noise)
        if rirs:
            self.plan.register_rirs(rirs)
        self.n_rir = len(rirs) if rirs else 0
        self.gen = torch.Generator()
        if seed is not None:
            self.gen.manual_seed(seed)

    def draw(self, B: int) -> AugParams:
        """One host-side draw per clip: apply flags, bank indices, noise offset, SNR."""
        g = self.gen
        p = AugParams()
        if self.n_rir:
            on = torch.rand(B, generator=g) < self.rir_prob
            pick = torch.randint(self.n_rir, (B,), generator=g, dtype=torch.int32)
            p.rir_idx = torch.where(on, pick, torch.full_like(pick, -1))
        if self.noise_lens:
            lo, hi = self.noise_snr_range
            on = torch.rand(B, generator=g) < self.background_noise_prob
            pick = torch.randint(len(self.noise_lens), (B,), generator=g, dtype=torch.int32)
            lens = torch.tensor(self.noise_lens, dtype=torch.float64)[pick.long()]
            p.noise_idx = torch.where(on, pick, torch.full_like(pick, -1))
            p.noise_off = (torch.rand(B, generator=g, dtype=torch.float64) * lens).long()
            p.snr_db = (lo + (hi - lo) * torch.rand(B, generator=g)).float()
        if self.time_stretch_prob > 0:
            lo, hi = self.time_stretch_range
            on = torch.rand(B, generator=g) < self.time_stretch_prob
            rate = lo + (hi - lo) * torch.rand(B, generator=g, dtype=torch.float64)
            p.stretch_rate = torch.where(on, rate, torch.ones_like(rate))
            p.stretch_lo = min(1.0, lo)
        if self.pitch_shift_prob > 0:
            lo, hi = self.pitch_shift_range
            on = torch.rand(B, generator=g) < self.pitch_shift_prob
            steps = torch.randint(lo, hi + 1, (B,), generator=g, dtype=torch.int32)       # random.randint: inclusive
            p.pitch_steps = torch.where(on, steps, torch.zeros_like(steps))
            p.pitch_range = (min(0, lo), max(0, hi))
        return p

    @torch.no_grad()
    def __call__(self, waveform: torch.Tensor, params: Optional[AugParams] = None) -> torch.Tensor:
        """(1, N) / (N,) / (B, N) -> same shape, on the input's device (tests/...:239-243)."""
        shape, src_dev = waveform.shape, waveform.device
        wav = waveform.reshape(-1, shape[-1])
        if params is None:
            params = self.draw(wav.shape[0])
        out = self.plan.augment(wav, params)
        return out.reshape(shape).to(src_dev)


class SpecAugment:
    def __init__(self, freq_mask_param: int = 15, time_mask_param: int = 35, n_freq_masks: int = 2,
                 n_time_masks: int = 2, mask_value: float = 0.0, seed: Optional[int] = None, device: str = "cuda"):
        self.freq_mask_param, self.time_mask_param = int(freq_mask_param), int(time_mask_param)
        self.n_freq_masks, self.n_time_masks = int(n_freq_masks), int(n_time_masks)
        self.mask_value = float(mask_value)
        self.device = torch.device(device)
        self.gen = torch.Generator()
        if seed is not None:
            self.gen.manual_seed(seed)

    def draw(self, B: int, n_feat: int, n_frames: int) -> AugParams:
        fs, fl = draw_mask_params(self.gen, B, n_feat, self.freq_mask_param, self.n_freq_masks)
        ts, tl = draw_mask_params(self.gen, B, n_frames, self.time_mask_param, self.n_time_masks)
        return AugParams(fmask_start=fs, fmask_len=fl, tmask_start=ts, tmask_len=tl)

    @torch.no_grad()
    def __call__(self, spec: torch.Tensor, params: Optional[AugParams] = None) -> torch.Tensor:
        """(..., F, T) -> same shape (tests/test_training_pipeline.py:259-262); the input is not modified."""
        if spec.dim() < 2:
            raise ValueError("spec must have at least (F, T) dimensions")
        shape, src_dev = spec.shape, spec.device
        dev = src_dev if src_dev.type == "cuda" else self.device
        dt = spec.dtype if spec.dtype in (torch.float32, torch.float16) else torch.float32
        x = spec.to(device=dev, dtype=dt).reshape(-1, shape[-2], shape[-1]).clone().contiguous()
        if params is None:
            params = self.draw(x.shape[0], shape[-2], shape[-1])
        spec_augment_(x, params.fmask_start, params.fmask_len, params.tmask_start, params.tmask_len, self.mask_value)
        return x.reshape(shape).to(device=src_dev, dtype=spec.dtype)
