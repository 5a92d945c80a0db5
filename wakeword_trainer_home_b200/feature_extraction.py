"""``FeatureExtractor`` - drop-in for the reference's ``src.data.feature_extraction.FeatureExtractor``.

Call surface reconstructed from its call sites (the module itself was never committed to the
reference): constructor kwargs ``sample_rate, feature_type, n_mels, n_mfcc,
n_fft, hop_length, device`` (src/evaluation/evaluator.py:86-94, src/evaluation/inference.py:94-102);
``__call__(waveform)`` takes a float tensor ``(N,)`` (evaluator.py:122-125) or ``(1, N)`` and
returns ``(1, F, T)`` with ``T = N // hop_length + 1`` (src/export/onnx_exporter.py:316-320);
callers then ``unsqueeze(0).to(device)`` it (evaluator.py:128).  Batches ``(B, N)`` /
``(B, 1, N)`` return ``(B, 1, F, T)``.
"""
from __future__ import annotations

import torch

from .pipeline import FeaturePlan


class FeatureExtractor:
    def __init__(self, sample_rate: int = 16000, feature_type: str = "mel", n_mels: int = 128, n_mfcc: int = 40,
                 n_fft: int = 1024, hop_length: int = 160, device: str = "cuda", **plan_kwargs):
        self.sample_rate, self.feature_type = sample_rate, feature_type
        self.n_mels, self.n_mfcc, self.n_fft, self.hop_length = n_mels, n_mfcc, n_fft, hop_length
        self.plan = FeaturePlan(sample_rate=sample_rate, feature_type=feature_type, n_mels=n_mels, n_mfcc=n_mfcc,
                                n_fft=n_fft, hop_length=hop_length, device=device, **plan_kwargs)
        self.device = self.plan.device

    @property
    def n_features(self) -> int:
        return self.plan.n_feat

    def output_shape(self, n_samples: int):
        return (1, self.plan.n_feat, self.plan.num_frames(n_samples))

    @torch.no_grad()
    def __call__(self, waveform: torch.Tensor) -> torch.Tensor:
        if not torch.is_tensor(waveform):
            waveform = torch.as_tensor(waveform)
        if waveform.dim() == 0 or waveform.dim() > 3 or (waveform.dim() == 3 and waveform.shape[1] != 1):
            raise ValueError(f"waveform must be (N,), (1, N), (B, N) or (B, 1, N); got {tuple(waveform.shape)}")
        single = waveform.dim() == 1 or (waveform.dim() == 2 and waveform.shape[0] == 1)
        wav = waveform.reshape(-1, waveform.shape[-1])
        feats = self.plan.featurize(wav)                 # (B, 1, F, T)
        return feats[0] if single else feats
