"""``WakewordDataset`` and ``load_dataset_splits`` - drop-ins for the reference's ``src.data.dataset`` (module absent
from its checkout; surface reconstructed from the call sites):

* ``WakewordDataset(manifest_path, sample_rate, audio_duration, augment, device, feature_type, n_mels, n_mfcc, n_fft,
  hop_length)``                                                   src/ui/panel_evaluation.py:417-428
* ``len(ds)``; ``ds[i] -> (features, label, metadata)`` with ``features`` a (1, F, T) float tensor that
  ``torch.stack`` + ``DataLoader(pin_memory=True)`` can batch, ``label`` an int, ``metadata`` a dict that may carry
  ``'path'``                                                      src/evaluation/evaluator.py:257-276,313
* ``load_dataset_splits(splits_dir, sample_rate, audio_duration, augment_train, augmentation_config, data_root, device,
  feature_type, n_mels, n_mfcc, n_fft, hop_length) -> (train_ds, val_ds, test_ds)``; reads ``train.json`` /
  ``val.json`` / ``test.json``; ``augmentation_config`` is the dict built at src/ui/panel_training.py:309-318;
  noise / RIR recordings live under ``data_root/raw/{background,rirs}`` (README.md:57-58)

Two ways to consume a dataset:

1. **The reference's way, unchanged**: hand it to ``torch.utils.data.DataLoader`` (panel_training.py:342-358,
   evaluator.py:270-277).  ``__getitem__`` then runs the GPU path on one clip and returns CPU tensors (the reference
   pins them).  With ``num_workers=0`` sequential reads are served from a chunk that was featurized in ONE launch
   (``chunk`` clips ahead), so the evaluator's own loop already gets batched GPU work.
2. **The fast way**: ``ds.loader(batch_size, ...)`` returns a ``DeviceBatchLoader`` / ``GpuBatchLoader`` that yields
   ``(inputs, targets)`` on the GPU - the iterable ``Trainer.train_epoch`` accepts (src/training/trainer.py:147-166) -
   with one fused pass per batch and augmentation draws made on the GPU.

Decoded clips are kept as int16 PCM (the files' native precision; half the bytes) in one (n, N) bank: on the GPU when
``resident=True`` (default when it fits ``resident_budget_bytes``), else pinned on the host.
"""
from __future__ import annotations

import os
from pathlib import Path
from typing import Any, Dict, List, Optional, Sequence, Tuple

import torch
from torch.utils.data import Dataset

from .config_adapter import normalize_feature_type
from .formats import load_split_manifest

AUDIO_EXTENSIONS = (".wav", ".wave", ".flac", ".mp3", ".ogg")


def _list_audio(folder) -> List[str]:
    folder = Path(folder)
    if not folder.is_dir():
        return []
    return sorted(str(p) for p in folder.rglob("*") if p.suffix.lower() in AUDIO_EXTENSIONS)


class WakewordDataset(Dataset):
    def __init__(self, manifest_path, sample_rate: int = 16000, audio_duration: float = 2.5, augment: bool = False,
                 device: str = "cuda", feature_type: str = "mel", n_mels: int = 128, n_mfcc: int = 40, n_fft: int = 1024,
                 hop_length: int = 160, augmentation_config: Optional[Dict[str, Any]] = None, data_root=None,
                 background_noise: Optional[Sequence[torch.Tensor]] = None, rirs: Optional[Sequence[torch.Tensor]] = None,
                 normalize_audio: bool = True, spec_augment: bool = False, seed: int = 0, chunk: int = 64,
                 resident: Optional[bool] = None, resident_budget_bytes: int = 8 << 30, return_device: str = "cpu"):
        self.manifest_path = Path(manifest_path)
        self.paths, self.labels, self.metadata = load_split_manifest(self.manifest_path)
        base = self.manifest_path.parent
        root = Path(data_root) if data_root is not None else None
        self.files = [self._resolve(p, base, root) for p in self.paths]
        self.sample_rate, self.audio_duration = int(sample_rate), float(audio_duration)
        self.n_samples = int(self.sample_rate * self.audio_duration)
        self.augment, self.device_arg = bool(augment), device
        self.feature_kw = dict(sample_rate=self.sample_rate, feature_type=normalize_feature_type(feature_type), n_mels=n_mels,
                               n_mfcc=n_mfcc, n_fft=n_fft, hop_length=hop_length)
        self.augmentation_config = dict(augmentation_config or {})
        self.data_root = root
        self._noise, self._rirs = background_noise, rirs
        self.normalize_audio, self.spec_augment, self.seed = bool(normalize_audio), bool(spec_augment), int(seed)
        self.chunk, self.return_device = max(1, int(chunk)), torch.device(return_device)
        self.resident, self.resident_budget_bytes = resident, int(resident_budget_bytes)
        self.epoch = 0
        # built lazily on first use (constructing a dataset needs no GPU; reading from it does)
        self._plan = None
        self._bank = None                    # (n, N) int16 PCM, CUDA or pinned host
        self._draw = None
        self._chunk_start, self._chunk_feats = -1, None

    # ------------------------------------------------------------------ manifest
    @staticmethod
    def _resolve(p: str, base: Path, root: Optional[Path]) -> str:
        if os.path.isabs(p) or os.path.exists(p):
            return p
        for b in ([root] if root is not None else []) + [base, base.parent, base.parent.parent]:
            if (b / p).exists():
                return str(b / p)
        return p

    def __len__(self) -> int:
        return len(self.files)

    def set_epoch(self, epoch: int):
        """New epoch = new augmentation draws for the same sample index (and a dropped read-ahead chunk)."""
        self.epoch = int(epoch)
        self._chunk_start = -1

    # ------------------------------------------------------------------ lazy GPU state
    @property
    def plan(self):
        if self._plan is None:
            from .pipeline import DrawConfig, FeaturePlan
            masks = dict(n_freq_masks=2, n_time_masks=2) if (self.augment and self.spec_augment) else {}
            self._plan = FeaturePlan(device=self.device_arg, **self.feature_kw, **masks)
            if self.augment:
                noise, rirs = self._load_banks()
                if noise:
                    self._plan.register_noise(noise)
                if rirs:
                    self._plan.register_rirs(rirs)
                c = self.augmentation_config
                snr = tuple(c.get("noise_snr_range", (5.0, 20.0)))
                self._draw = DrawConfig(seed=self.seed, rir_prob=float(c.get("rir_prob", 0.25)),
                                        noise_prob=float(c.get("background_noise_prob", 0.5)),
                                        freq_mask_prob=float(c.get("freq_mask_prob", 0.5)) if self.spec_augment else 0.0,
                                        time_mask_prob=float(c.get("time_mask_prob", 0.5)) if self.spec_augment else 0.0,
                                        snr_range=(float(snr[0]), float(snr[1])),
                                        stretch_prob=float(c.get("time_stretch_prob", 0.0)),
                                        stretch_range=tuple(c.get("time_stretch_range", (0.8, 1.2))),
                                        pitch_prob=float(c.get("pitch_shift_prob", 0.0)),
                                        pitch_range=tuple(int(v) for v in c.get("pitch_shift_range", (-2, 2))))
        return self._plan

    def _load_banks(self):
        noise, rirs = self._noise, self._rirs
        if (noise is None or rirs is None) and self.data_root is not None:
            from .audio_utils import AudioProcessor
            ap = AudioProcessor(self.sample_rate, None, plan=self._plan, normalize=False)
            if noise is None:
                noise = [torch.from_numpy(ap.process_audio(f)) for f in _list_audio(self.data_root / "raw" / "background")]
            if rirs is None:
                rirs = [torch.from_numpy(ap.process_audio(f))[:16384] for f in _list_audio(self.data_root / "raw" / "rirs")]
        return list(noise or []), list(rirs or [])

    @property
    def bank(self) -> torch.Tensor:
        """All clips decoded, resampled, normalised and padded / trimmed once, kept as (n, N) int16 PCM."""
        if self._bank is None:
            from .audio_utils import AudioProcessor
            plan = self.plan
            ap = AudioProcessor(self.sample_rate, self.audio_duration, plan=plan, normalize=self.normalize_audio)
            n, N = len(self.files), self.n_samples
            on_gpu = self.resident if self.resident is not None else (2 * n * N <= self.resident_budget_bytes)
            bank = torch.empty(n, N, dtype=torch.int16, device=plan.device) if on_gpu else torch.empty(n, N, dtype=torch.int16).pin_memory()
            for a in range(0, n, 256):
                x = ap.process_batch(self.files[a:a + 256])[:, :N]
                pcm = (x.clamp(-1.0, 32767.0 / 32768.0) * 32768.0).round().to(torch.int16)
                bank[a:a + 256].copy_(pcm, non_blocking=True)
            torch.cuda.synchronize(plan.device)
            self._bank = bank
        return self._bank

    # ------------------------------------------------------------------ the reference's per-item access
    def _featurize_rows(self, rows: torch.Tensor) -> torch.Tensor:
        from .pipeline import gather_clips
        plan, bank = self.plan, self.bank
        if bank.is_cuda:
            wav = gather_clips(bank, rows.to(bank.device))
        else:
            wav = bank.index_select(0, rows).pin_memory().to(plan.device, non_blocking=True).float() / 32768.0
        aug = None
        if self.augment and self._draw is not None:
            # counter-based draws: sample number = epoch * len + index, so an item's augmentation is reproducible
            first = self.epoch * len(self) + int(rows[0])
            aug = plan.draw_aug(self._draw, first, rows.numel(), self.n_samples)
        return plan.featurize(wav, aug)

    def __getitem__(self, i: int) -> Tuple[torch.Tensor, int, dict]:
        n = len(self)
        if i < 0:
            i += n
        if not 0 <= i < n:
            raise IndexError(i)
        if not (self._chunk_start <= i < self._chunk_start + (0 if self._chunk_feats is None else self._chunk_feats.shape[0])):
            rows = torch.arange(i, min(i + self.chunk, n), dtype=torch.int64)
            feats = self._featurize_rows(rows)
            self._chunk_feats = feats if self.return_device.type == "cuda" else feats.cpu()
            self._chunk_start = i
        f = self._chunk_feats[i - self._chunk_start]
        meta = dict(self.metadata[i])
        meta.setdefault("path", self.paths[i])
        return f, int(self.labels[i]), meta

    # ------------------------------------------------------------------ the fast way
    def loader(self, batch_size: int, shuffle: bool = True, drop_last: bool = False, rank: int = 0, world_size: int = 1,
               seed: Optional[int] = None, with_metadata: bool = False):
        """Batched GPU loader over this dataset: yields ``(inputs, targets)`` (or ``(inputs, targets, metadata)``) like
        the DataLoader the reference's Trainer / evaluator iterate, one fused pass per batch."""
        from .loader import DeviceBatchLoader, GpuBatchLoader
        plan, bank = self.plan, self.bank
        seed = self.seed if seed is None else seed
        if bank.is_cuda and not with_metadata:
            return DeviceBatchLoader(bank, self.labels, plan, batch_size, draw=self._draw if self.augment else None, shuffle=shuffle,
                                     seed=seed, rank=rank, world_size=world_size, drop_last=drop_last)
        clips = bank.float().div_(32768.0) if not bank.is_cuda else bank.float().div_(32768.0)
        return GpuBatchLoader(clips, self.labels, plan, batch_size, shuffle=shuffle, seed=seed, rank=rank, world_size=world_size,
                              drop_last=drop_last, metadata=[dict(m, path=p) for m, p in zip(self.metadata, self.paths)] if with_metadata else None)


def load_dataset_splits(splits_dir, sample_rate: int = 16000, audio_duration: float = 2.5, augment_train: bool = True,
                        augmentation_config: Optional[Dict[str, Any]] = None, data_root=None, device: str = "cuda",
                        feature_type: str = "mel", n_mels: int = 128, n_mfcc: int = 40, n_fft: int = 1024,
                        hop_length: int = 160, **dataset_kwargs) -> Tuple[WakewordDataset, WakewordDataset, WakewordDataset]:
    """(train, val, test) datasets from ``splits_dir/{train,val,test}.json`` (src/ui/panel_training.py:323-336).
    Only the training split is augmented (``augment_train``)."""
    splits_dir = Path(splits_dir)
    common = dict(sample_rate=sample_rate, audio_duration=audio_duration, device=device, feature_type=feature_type,
                  n_mels=n_mels, n_mfcc=n_mfcc, n_fft=n_fft, hop_length=hop_length, data_root=data_root, **dataset_kwargs)
    out = []
    for name, aug in (("train", augment_train), ("val", False), ("test", False)):
        path = splits_dir / f"{name}.json"
        if not path.exists():
            raise FileNotFoundError(f"Dataset split not found: {path}")
        out.append(WakewordDataset(path, augment=aug, augmentation_config=augmentation_config if aug else None, **common))
    return tuple(out)
