"""Host-fed streaming front end of the feature path and the batched loader that replaces the
reference's ``WakewordDataset`` + ``DataLoader`` pair.

``StreamedFeaturizer`` keeps the GPU busy when clips live in HOST memory: a copy stream uploads
batch i+1 (pinned clips + augmentation draws) while the compute stream featurizes batch i and a
third stream downloads the features of batch i-1.  It is plain use of the public
``FeaturePlan.featurize`` call on torch streams - the end-to-end path bench.py times.

``GpuBatchLoader`` iterates ``(inputs, targets)`` exactly like the object the reference's
``Trainer.train_epoch`` consumes (src/training/trainer.py:147-157): one fused GPU pass per batch
instead of 16 DataLoader workers calling torchaudio per sample (src/ui/panel_training.py:342-349).
It shards clip indices across ranks (no collective) and redraws augmentations from
``shard_seed(seed, rank, step)`` so a resumed run reproduces them.
"""
from __future__ import annotations

from typing import Iterator, Optional, Tuple

import torch

from .pipeline import AugParams, DrawConfig, FeaturePlan, gather_clips
from .sharding import shard_range, shard_seed


class StreamedFeaturizer:
    """Double-buffered host -> device -> host featurization through ``FeaturePlan.featurize``."""

    def __init__(self, plan: FeaturePlan, batch: int, n_samples: int, depth: int = 2, copy_back: bool = True,
                 pcm16: bool = False):
        """pcm16=True: host batches are int16 PCM (the native format of WAV files); they cross PCIe at half
        the bytes and are converted on the device (x / 32768, exact) by ``wwf_gather_clips``."""
        self.plan, self.depth, self.copy_back, self.pcm16 = plan, depth, copy_back, pcm16
        dev = plan.device
        T = plan.num_frames(n_samples)
        self.s_in, self.s_run, self.s_out = (torch.cuda.Stream(dev) for _ in range(3))
        self.d_wav = [torch.empty(batch, n_samples, dtype=torch.float32, device=dev) for _ in range(depth)]
        self.d_pcm = [torch.empty(batch, n_samples, dtype=torch.int16, device=dev) for _ in range(depth)] if pcm16 else None
        self._rows = torch.arange(batch, dtype=torch.int64, device=dev)
        self.d_out = [torch.empty(batch, 1, plan.n_feat, T, dtype=plan.out_dtype, device=dev) for _ in range(depth)]
        self.h_out = [torch.empty(batch, 1, plan.n_feat, T, dtype=plan.out_dtype).pin_memory() for _ in range(depth)] if copy_back else None
        self.ev_in = [torch.cuda.Event() for _ in range(depth)]
        self.ev_run = [torch.cuda.Event() for _ in range(depth)]
        self.ev_out = [torch.cuda.Event() for _ in range(depth)]
        self.n = 0

    def submit(self, host_wav: torch.Tensor, host_aug: Optional[AugParams]) -> int:
        """Queue one batch (pinned host clips + host draws).  Returns its slot; the features are
        in ``d_out[slot]`` after ``ev_run[slot]`` and in ``h_out[slot]`` after ``ev_out[slot]``."""
        k = self.n % self.depth
        self.n += 1
        with torch.cuda.stream(self.s_in):
            self.s_in.wait_event(self.ev_run[k])          # previous use of this input slot has been consumed
            (self.d_pcm if self.pcm16 else self.d_wav)[k].copy_(host_wav, non_blocking=True)
            aug = None if host_aug is None else host_aug.to(self.plan.device, non_blocking=True)
            self.ev_in[k].record(self.s_in)
        with torch.cuda.stream(self.s_run):
            self.s_run.wait_event(self.ev_in[k])
            self.s_run.wait_event(self.ev_out[k])         # previous features of this slot have left
            if self.pcm16:
                gather_clips(self.d_pcm[k], self._rows, out=self.d_wav[k])
            self.plan.featurize(self.d_wav[k], aug, out=self.d_out[k])
            if aug is not None:                           # keep the draw tensors alive until the kernel ran
                for f in aug.tensor_fields():
                    v = getattr(aug, f)
                    if v is not None:
                        v.record_stream(self.s_run)
            self.ev_run[k].record(self.s_run)
        if self.copy_back:
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_event(self.ev_run[k])
                self.h_out[k].copy_(self.d_out[k], non_blocking=True)
                self.ev_out[k].record(self.s_out)
        return k

    def wait(self, slot: int) -> torch.Tensor:
        (self.ev_out if self.copy_back else self.ev_run)[slot].synchronize()
        return self.h_out[slot] if self.copy_back else self.d_out[slot]

    def synchronize(self):
        for s in (self.s_in, self.s_run, self.s_out):
            s.synchronize()


class GpuBatchLoader:
    """``for inputs, targets in loader`` over an in-memory clip set, features computed on the GPU.

    clips    (n, N) float32, pinned host or device-resident
    labels   (n,) integer targets
    augment  optional ``AudioAugmentation``-like object with ``draw(B) -> AugParams``
    spec_augment optional ``SpecAugment``-like object with ``draw(B, F, T) -> AugParams``;
             its ``n_freq_masks / n_time_masks`` must match the plan's
    metadata optional sequence of per-clip dicts; batches then carry a third element, the list of the
             batch's dicts (the evaluator's ``(features, labels, metadata_list)`` batches)
    """

    def __init__(self, clips: torch.Tensor, labels: torch.Tensor, plan: FeaturePlan, batch_size: int,
                 augment=None, spec_augment=None, shuffle: bool = True, seed: int = 0, rank: int = 0,
                 world_size: int = 1, drop_last: bool = False, metadata=None):
        if clips.dim() != 2 or labels.shape[0] != clips.shape[0]:
            raise ValueError("clips must be (n, N) and labels (n,)")
        self.clips, self.labels, self.plan = clips, labels, plan
        self.batch_size, self.shuffle, self.seed = batch_size, shuffle, seed
        self.rank, self.world, self.drop_last = rank, world_size, drop_last
        self.augment, self.spec_augment = augment, spec_augment
        # per-sample metadata dicts (e.g. {'path': ...}): when given, batches are (inputs, targets, [meta, ...]) like
        # the collate function of the reference's evaluator (src/evaluation/evaluator.py:257-268, :313)
        self.metadata = metadata
        self.epoch = 0
        self.step = 0                   # batches produced in the whole run: keys the augmentation draws
        self._skip = 0                  # batches of the current epoch to skip after load_state_dict (mid-epoch resume)
        self._batches_done = 0          # batches produced so far in the current epoch

    def set_epoch(self, epoch: int):
        self.epoch = epoch

    # ---- resume: sits beside the reference's checkpoint dict (src/training/trainer.py:486-525) ----
    def state_dict(self) -> dict:
        """Everything a resumed run needs to see the same batches with the same augmentation draws: the run seed,
        the epoch (it fixes the permutation) and the number of batches already produced (it keys the per-step draws)."""
        return {"kind": "GpuBatchLoader", "seed": self.seed, "epoch": self.epoch, "step": self.step,
                "batches_done": self._batches_done, "rank": self.rank, "world_size": self.world}

    def load_state_dict(self, state: dict):
        if state.get("kind") != "GpuBatchLoader":
            raise ValueError(f"not a GpuBatchLoader state: {state.get('kind')!r}")
        if (state["rank"], state["world_size"]) != (self.rank, self.world):
            raise ValueError("loader state was saved for another rank / world size")
        self.seed, self.epoch, self.step = int(state["seed"]), int(state["epoch"]), int(state["step"])
        self._skip = int(state.get("batches_done", 0))

    def _indices(self) -> torch.Tensor:
        n = self.clips.shape[0]
        if self.shuffle:
            g = torch.Generator().manual_seed(self.seed + self.epoch)     # same permutation on every rank
            perm = torch.randperm(n, generator=g)
        else:
            perm = torch.arange(n)
        a, b = shard_range(n, self.rank, self.world)
        return perm[a:b]

    def __len__(self) -> int:
        a, b = shard_range(self.clips.shape[0], self.rank, self.world)
        n = b - a
        return n // self.batch_size if self.drop_last else (n + self.batch_size - 1) // self.batch_size

    def __iter__(self) -> Iterator[Tuple[torch.Tensor, torch.Tensor]]:
        idx = self._indices()
        dev = self.plan.device
        on_dev = self.clips.is_cuda
        first, self._skip = self._skip, 0
        if first >= len(self):                                            # the saved epoch was complete
            first = 0
        self._batches_done = first
        for i in range(first, len(self)):
            sel = idx[i * self.batch_size:(i + 1) * self.batch_size]
            B = sel.numel()
            wav = self.clips.index_select(0, sel.to(self.clips.device))
            if not on_dev:
                wav = wav.pin_memory().to(dev, non_blocking=True)
            g = torch.Generator().manual_seed(shard_seed(self.seed, self.rank, self.step))
            aug = None
            if self.augment is not None:
                self.augment.gen = g
                aug = self.augment.draw(B)
            if self.spec_augment is not None:
                self.spec_augment.gen = g
                m = self.spec_augment.draw(B, self.plan.n_feat, self.plan.num_frames(wav.shape[1]))
                aug = aug or AugParams()
                aug.fmask_start, aug.fmask_len, aug.tmask_start, aug.tmask_len = m.fmask_start, m.fmask_len, m.tmask_start, m.tmask_len
            self.step += 1
            self._batches_done = i + 1
            feats = self.plan.featurize(wav, aug)
            targets = self.labels.index_select(0, sel.to(self.labels.device)).to(dev, non_blocking=True)
            if self.metadata is not None:
                yield feats, targets, [self.metadata[i] for i in sel.tolist()]
            else:
                yield feats, targets


class DeviceBatchLoader:
    """Fully device-resident variant of ``GpuBatchLoader``: the clip bank (float32 or int16 PCM) lives in
    HBM, batches are assembled by ``wwf_gather_clips`` and the augmentation draws are made on the GPU by
    ``wwf_draw_aug`` (counter = number of the sample in the run, so a resumed run that restores
    ``samples_drawn`` reproduces them).  A training step issues no host->device copy at all.
    Yields ``(inputs, targets)`` like the DataLoader ``Trainer.train_epoch`` iterates
    (src/training/trainer.py:147-157)."""

    def __init__(self, bank: torch.Tensor, labels: torch.Tensor, plan: FeaturePlan, batch_size: int,
                 draw: Optional[DrawConfig] = None, shuffle: bool = True, seed: int = 0, rank: int = 0,
                 world_size: int = 1, drop_last: bool = False):
        if not bank.is_cuda or bank.dim() != 2 or labels.shape[0] != bank.shape[0]:
            raise ValueError("bank must be a CUDA (n, N) tensor and labels (n,)")
        self.bank, self.plan, self.batch_size = bank, plan, batch_size
        self.labels = labels.to(bank.device)
        self.draw, self.shuffle, self.seed = draw, shuffle, seed
        self.rank, self.world, self.drop_last = rank, world_size, drop_last
        self.epoch = 0
        self.samples_drawn = 0          # checkpoint this (with seed) to resume with identical augmentations
        self._skip = 0                  # batches of the current epoch to skip after load_state_dict (mid-epoch resume)
        self._batches_done = 0          # batches produced so far in the current epoch

    def set_epoch(self, epoch: int):
        self.epoch = epoch

    # ---- resume: sits beside the reference's checkpoint dict (src/training/trainer.py:486-525) ----
    def state_dict(self) -> dict:
        """seed + epoch fix the permutation, ``samples_drawn`` is the Philox counter of the next augmentation draw and
        ``batches_done`` the position inside the epoch: a loader restored from this continues with bit-identical batches."""
        return {"kind": "DeviceBatchLoader", "seed": self.seed, "epoch": self.epoch, "samples_drawn": self.samples_drawn,
                "batches_done": self._batches_done, "rank": self.rank, "world_size": self.world}

    def load_state_dict(self, state: dict):
        if state.get("kind") != "DeviceBatchLoader":
            raise ValueError(f"not a DeviceBatchLoader state: {state.get('kind')!r}")
        if (state["rank"], state["world_size"]) != (self.rank, self.world):
            raise ValueError("loader state was saved for another rank / world size")
        self.seed, self.epoch = int(state["seed"]), int(state["epoch"])
        self.samples_drawn = int(state["samples_drawn"])
        self._skip = int(state.get("batches_done", 0))

    def __len__(self) -> int:
        a, b = shard_range(self.bank.shape[0], self.rank, self.world)
        n = b - a
        return n // self.batch_size if self.drop_last else (n + self.batch_size - 1) // self.batch_size

    def __iter__(self) -> Iterator[Tuple[torch.Tensor, torch.Tensor]]:
        n = self.bank.shape[0]
        if self.shuffle:
            g = torch.Generator().manual_seed(self.seed + self.epoch)     # same permutation on every rank
            perm = torch.randperm(n, generator=g)
        else:
            perm = torch.arange(n)
        a, b = shard_range(n, self.rank, self.world)
        idx = perm[a:b].to(self.bank.device)                              # one small upload per epoch
        # ranks draw from disjoint counter ranges: sample numbers are offset by rank * 2^40
        base = (self.rank << 40) + self.samples_drawn
        first, self._skip = self._skip, 0
        if first >= len(self):                                            # the saved epoch was complete
            first = 0
        self._batches_done = first
        for i in range(first, len(self)):
            sel = idx[i * self.batch_size:(i + 1) * self.batch_size]
            wav = gather_clips(self.bank, sel)
            aug = None
            if self.draw is not None:
                aug = self.plan.draw_aug(self.draw, base, sel.numel(), wav.shape[1])
            base += sel.numel()
            self.samples_drawn += sel.numel()
            self._batches_done = i + 1
            yield self.plan.featurize(wav, aug), self.labels.index_select(0, sel)
