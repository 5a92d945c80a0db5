"""On-disk formats either side of the feature path.

The reference documents three ``.npy`` shapes (src/ui/panel_docs.py:124-130): raw audio ``(N, samples)``,
spectrograms ``(N, freq_bins, time_steps)`` and MFCC ``(N, n_mfcc, time_steps)``, and keeps its splits as JSON
manifests ``data/splits/{train,val,test}.json`` / ``dataset_manifest.json`` (src/ui/panel_dataset.py:206,278-279,
src/ui/panel_training.py:299-300, src/ui/panel_evaluation.py:407).  The code that writes those manifests
(``src/data/splitter.py``) is absent from the reference checkout, so their schema is not knowable; the reader below
accepts the shapes such a file can reasonably have - a list of entries or a dict holding one under ``files`` /
``samples`` / ``items`` / ``data`` - where an entry is a path string or a dict with ``path`` (the key the evaluator
reads back from the sample metadata, src/evaluation/evaluator.py:313) and ``label`` or ``category``.
"""
from __future__ import annotations

import json
import os
from typing import Iterable, List, Optional, Sequence, Tuple

import numpy as np
import torch

from .pipeline import AugParams, FeaturePlan

POSITIVE_CATEGORIES = ("positive", "wakeword", "wake_word", "1", "true")


def _entry(e) -> Tuple[str, int, dict]:
    if isinstance(e, str):
        return e, -1, {"path": e}
    if not isinstance(e, dict) or "path" not in e:
        raise ValueError(f"manifest entry without a 'path': {e!r}")
    if "label" in e:
        label = int(e["label"])
    elif "category" in e:
        label = int(str(e["category"]).lower() in POSITIVE_CATEGORIES)
    else:
        label = -1
    return str(e["path"]), label, dict(e)


def load_split_manifest(path) -> Tuple[List[str], torch.Tensor, List[dict]]:
    """``data/splits/*.json`` -> (paths, int64 labels (-1 = unknown), per-sample metadata dicts)."""
    with open(path) as f:
        doc = json.load(f)
    items = doc
    if isinstance(doc, dict):
        for key in ("files", "samples", "items", "data"):
            if isinstance(doc.get(key), list):
                items = doc[key]
                break
        else:
            raise ValueError(f"{path}: no list of samples found (keys: {sorted(doc)})")
    rows = [_entry(e) for e in items]
    return [r[0] for r in rows], torch.tensor([r[1] for r in rows], dtype=torch.int64), [r[2] for r in rows]


def save_split_manifest(path, paths: Sequence[str], labels: Iterable[int], extra: Optional[dict] = None) -> None:
    """Write a manifest in the shape ``load_split_manifest`` (and a ``meta['path']`` reader) understands."""
    doc = dict(extra or {})
    doc["files"] = [{"path": str(p), "label": int(l), "category": "positive" if int(l) == 1 else "negative"}
                    for p, l in zip(paths, labels)]
    os.makedirs(os.path.dirname(os.path.abspath(path)) or ".", exist_ok=True)
    with open(path, "w") as f:
        json.dump(doc, f, indent=2)


def npy_kind(arr: np.ndarray) -> str:
    """'audio' for (N, samples), 'features' for (N, freq_bins | n_mfcc, time_steps) (panel_docs.py:124-130)."""
    if arr.ndim == 2:
        return "audio"
    if arr.ndim == 3 or (arr.ndim == 4 and arr.shape[1] == 1):
        return "features"
    raise ValueError(f"unsupported .npy shape {arr.shape}: expected (N, samples) or (N, F, T)")


def load_npy(path, mmap: bool = True) -> Tuple[np.ndarray, str]:
    arr = np.load(path, mmap_mode="r" if mmap else None)
    return arr, npy_kind(arr)


@torch.no_grad()
def precompute_features(plan: FeaturePlan, clips, out_path=None, batch_size: int = 1024,
                        aug: Optional[AugParams] = None) -> np.ndarray:
    """(N, samples) raw audio (numpy, memory-mapped .npy, or tensor) -> (N, F, T) features computed on the GPU in
    batches and, when ``out_path`` is given, written as the reference's pre-computed-feature ``.npy``."""
    n, N = clips.shape
    T = plan.num_frames(N)
    np_dtype = np.float16 if plan.out_dtype == torch.float16 else np.float32
    if out_path is not None:
        os.makedirs(os.path.dirname(os.path.abspath(out_path)) or ".", exist_ok=True)
        out = np.lib.format.open_memmap(out_path, mode="w+", dtype=np_dtype, shape=(n, plan.n_feat, T))
    else:
        out = np.empty((n, plan.n_feat, T), np_dtype)
    for a in range(0, n, batch_size):
        b = min(a + batch_size, n)
        x = clips[a:b]
        x = torch.from_numpy(np.array(x, dtype=np.float32)) if isinstance(x, np.ndarray) else x    # copy: memory maps are read-only
        out[a:b] = plan.featurize(x.to(plan.device, torch.float32), aug)[:, 0].cpu().numpy()
    if out_path is not None:
        out.flush()
    return out


class NpyFeatureLoader:
    """Iterate ``(inputs, targets)`` over a pre-computed feature file (N, F, T): what ``Trainer.train_epoch``
    consumes (src/training/trainer.py:147-157) when features were cached offline.  Batches are gathered on the host
    from the memory map, uploaded through pinned memory and returned as (B, 1, F, T) CUDA tensors."""

    def __init__(self, features, labels, batch_size: int, device="cuda", shuffle: bool = True, seed: int = 0,
                 rank: int = 0, world_size: int = 1, drop_last: bool = False):
        from .sharding import shard_range
        self.features = features if not isinstance(features, (str, os.PathLike)) else np.load(features, mmap_mode="r")
        if npy_kind(self.features) != "features":
            raise ValueError("NpyFeatureLoader wants (N, F, T) features; use GpuBatchLoader for (N, samples) audio")
        self.labels = torch.as_tensor(labels)
        if self.labels.shape[0] != self.features.shape[0]:
            raise ValueError("labels and features disagree on N")
        self.batch_size, self.device, self.shuffle, self.seed = batch_size, torch.device(device), shuffle, seed
        self.drop_last, self.epoch = drop_last, 0
        self._range = shard_range(self.features.shape[0], rank, world_size)

    def set_epoch(self, epoch: int):
        self.epoch = epoch

    def __len__(self) -> int:
        n = self._range[1] - self._range[0]
        return n // self.batch_size if self.drop_last else (n + self.batch_size - 1) // self.batch_size

    def __iter__(self):
        n = self.features.shape[0]
        perm = torch.randperm(n, generator=torch.Generator().manual_seed(self.seed + self.epoch)) if self.shuffle else torch.arange(n)
        idx = perm[self._range[0]:self._range[1]]
        for i in range(len(self)):
            sel = idx[i * self.batch_size:(i + 1) * self.batch_size]
            order = torch.sort(sel).values.numpy()                    # ascending reads from the memory map
            x = torch.from_numpy(np.array(self.features[order])).pin_memory().to(self.device, non_blocking=True)
            if x.dim() == 3:
                x = x.unsqueeze(1)
            yield x, self.labels[torch.from_numpy(order)].to(self.device, non_blocking=True)
