"""``AudioProcessor`` - drop-in for the reference's ``src.data.audio_utils.AudioProcessor`` (SURVEY.md section 8f
row 3): the step in FRONT of the feature path.  The reference module is absent from its checkout; the surface
is reconstructed from the call sites:

* ``AudioProcessor(target_sr=sample_rate, target_duration=audio_duration)``  src/evaluation/evaluator.py:76-79,
  src/evaluation/inference.py:84-87
* ``process_audio(path) -> numpy float array`` that the caller turns into ``torch.from_numpy(audio).float()`` and
  hands to FeatureExtractor (evaluator.py:119-125, :183)
* accepted inputs: 8-48 kHz, 16-bit, mono or stereo ("stereo will be converted"), WAV / MP3 / FLAC / OGG
  (src/ui/panel_docs.py:134-138); ``DataConfig.normalize_audio`` (src/config/defaults.py:24).

What runs where: the container parsing stays on the host (stdlib ``wave`` / a RIFF reader for PCM and IEEE-float
WAV; other containers need an installed decoder and raise otherwise), everything numeric runs on the GPU:
mono mix, sample-rate conversion (``wwf_resample`` = torchaudio ``F.resample`` arithmetic), peak normalisation
(``wwf_peak_normalize``), pad / trim to ``target_duration``.
Choices the missing module leaves open, made explicit here: channels are averaged; a clip longer than the target
keeps its first ``target`` samples; a shorter one is zero-padded at the end.
"""
from __future__ import annotations

import os
import struct
from typing import Optional, Sequence, Tuple

import numpy as np
import torch

from .pipeline import FeaturePlan, peak_normalize


def read_wav(path: str) -> Tuple[np.ndarray, int]:
    """Decode a RIFF/WAVE file: PCM 8/16/24/32-bit or IEEE float 32/64 -> (float32 [channels][frames] in [-1, 1), rate)."""
    with open(path, "rb") as f:
        data = f.read()
    if len(data) < 12 or data[:4] != b"RIFF" or data[8:12] != b"WAVE":
        raise ValueError(f"{path}: not a RIFF/WAVE file")
    pos, fmt, pcm = 12, None, None
    while pos + 8 <= len(data):
        cid, size = data[pos:pos + 4], struct.unpack("<I", data[pos + 4:pos + 8])[0]
        body = data[pos + 8:pos + 8 + size]
        if cid == b"fmt ":
            fmt = struct.unpack("<HHIIHH", body[:16])
            if fmt[0] == 0xFFFE and len(body) >= 26:            # WAVE_FORMAT_EXTENSIBLE: real tag in the sub-format GUID
                fmt = (struct.unpack("<H", body[24:26])[0],) + fmt[1:]
        elif cid == b"data":
            pcm = body
        pos += 8 + size + (size & 1)
    if fmt is None or pcm is None:
        raise ValueError(f"{path}: missing fmt or data chunk")
    tag, ch, rate, _, _, bits = fmt
    if ch < 1:
        raise ValueError(f"{path}: {ch} channels")
    if tag == 1:                                               # integer PCM
        if bits == 8:
            x = (np.frombuffer(pcm, np.uint8).astype(np.float32) - 128.0) / 128.0
        elif bits == 16:
            x = np.frombuffer(pcm[:len(pcm) // 2 * 2], "<i2").astype(np.float32) / 32768.0
        elif bits == 24:
            b = np.frombuffer(pcm[:len(pcm) // 3 * 3], np.uint8).reshape(-1, 3).astype(np.int32)
            v = b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16)
            x = ((v ^ 0x800000) - 0x800000).astype(np.float32) / 8388608.0
        elif bits == 32:
            x = (np.frombuffer(pcm[:len(pcm) // 4 * 4], "<i4").astype(np.float64) / 2147483648.0).astype(np.float32)
        else:
            raise ValueError(f"{path}: unsupported PCM bit depth {bits}")
    elif tag == 3:                                             # IEEE float
        x = np.frombuffer(pcm[:len(pcm) // (bits // 8) * (bits // 8)], "<f4" if bits == 32 else "<f8").astype(np.float32)
    else:
        raise ValueError(f"{path}: unsupported WAVE format tag {tag}")
    frames = x.size // ch
    return np.ascontiguousarray(x[:frames * ch].reshape(frames, ch).T), int(rate)


def _decode(path: str) -> Tuple[np.ndarray, int]:
    if os.path.splitext(path)[1].lower() in (".wav", ".wave"):
        return read_wav(path)
    try:                                                       # MP3 / FLAC / OGG need a decoder the image may not have
        import soundfile as sf
    except Exception as e:  # pragma: no cover
        raise RuntimeError(f"{path}: only WAV is decoded natively; install soundfile for other containers") from e
    x, rate = sf.read(path, dtype="float32", always_2d=True)   # pragma: no cover
    return np.ascontiguousarray(x.T), int(rate)                # pragma: no cover


class AudioProcessor:
    def __init__(self, target_sr: int = 16000, target_duration: Optional[float] = None, device: str = "cuda",
                 normalize: bool = False, plan: Optional[FeaturePlan] = None):
        self.target_sr, self.target_duration, self.normalize = int(target_sr), target_duration, bool(normalize)
        self.target_samples = None if target_duration is None else int(round(target_duration * target_sr))
        # the resampler's coefficient tables hang off a plan; its feature settings are irrelevant here
        self.plan = plan if plan is not None else FeaturePlan(sample_rate=target_sr, n_fft=400, hop_length=160, n_mels=40, device=device)
        self.device = self.plan.device

    @torch.no_grad()
    def process_waveform(self, wav, sample_rate: int) -> torch.Tensor:
        """(channels, frames) or (frames,) at ``sample_rate`` -> (target_samples,) float32 on the GPU."""
        x = torch.as_tensor(wav, dtype=torch.float32).to(self.device)
        if x.dim() == 2:
            x = x.mean(dim=0)                                  # stereo -> mono
        if x.dim() != 1 or x.numel() == 0:
            raise ValueError(f"expected (channels, frames) or (frames,), got {tuple(x.shape)}")
        x = x.reshape(1, -1)
        if int(sample_rate) != self.target_sr:
            x = self.plan.resample(x, int(sample_rate), self.target_sr)
        if self.normalize:
            x = peak_normalize(x.contiguous())
        n = self.target_samples
        if n is not None:
            x = x[:, :n] if x.shape[1] >= n else torch.nn.functional.pad(x, (0, n - x.shape[1]))
        return x[0].contiguous()

    def process_audio(self, path) -> np.ndarray:
        """File -> mono float32 numpy array at target_sr, padded / trimmed to target_duration
        (what evaluator.py:119-122 passes to torch.from_numpy(...).float())."""
        wav, rate = _decode(str(path))
        return self.process_waveform(wav, rate).cpu().numpy()

    @torch.no_grad()
    def process_batch(self, paths: Sequence) -> torch.Tensor:
        """Many files -> (B, target_samples) float32 on the GPU (needs target_duration); files sharing a sample
        rate are resampled in one launch."""
        if self.target_samples is None:
            raise ValueError("process_batch needs target_duration")
        out = torch.zeros(len(paths), self.target_samples, dtype=torch.float32, device=self.device)
        by_rate: dict = {}
        for i, p in enumerate(paths):
            wav, rate = _decode(str(p))
            by_rate.setdefault(rate, []).append((i, torch.from_numpy(wav.mean(axis=0) if wav.shape[0] > 1 else wav[0])))
        for rate, items in by_rate.items():
            longest = max(w.numel() for _, w in items)
            batch = torch.zeros(len(items), longest)
            for r, (_, w) in enumerate(items):
                batch[r, :w.numel()] = w                        # trailing zeros resample to (near) zeros
            x = batch.to(self.device)
            if rate != self.target_sr:
                x = self.plan.resample(x, rate, self.target_sr)
            if self.normalize:
                # like process_waveform: the peak is taken over the WHOLE file, before the trim to target_duration
                # (rows are zero-padded to the longest file of the group, which does not move a peak)
                x = peak_normalize(x.contiguous())
            n = min(x.shape[1], self.target_samples)
            for r, (i, w) in enumerate(items):
                keep = min(n, -(-w.numel() * self.target_sr // rate))      # ceil(new * len / orig): this file's own length
                out[i, :keep] = x[r, :keep]
        return out
