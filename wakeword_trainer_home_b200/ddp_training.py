"""BASELINE.json configs[2]: the B200 feature path feeding the reference's OWN train step under DistributedDataParallel.

The reference trains with ``Trainer(model, train_loader, val_loader, config, checkpoint_dir, device)``
(src/training/trainer.py:49-121); ``train_epoch`` iterates any iterable of ``(inputs, targets)`` batches
(src/training/trainer.py:147-166) and the UI builds that iterable as a 16-worker ``DataLoader`` over
``WakewordDataset`` (src/ui/panel_training.py:323-358).  Here the iterable is ``DeviceBatchLoader``: clip bank in HBM,
batch gather, augmentation draws and features on the GPU, no host->device copy per step.  The classifier is the
reference's ``create_model(cfg.model.architecture)`` (src/models/architectures.py:437) - cuDNN, untouched - wrapped in
stock ``DistributedDataParallel``; its gradient all-reduce over NCCL / NVLink is the ONLY collective of the job (the
feature path shards by clip index and has none).  Nothing of the reference is modified: the
module is imported from /root/reference or baseline/_ref, with this package mounted as its missing ``src.data``.
"""
from __future__ import annotations

import os
import tempfile
from pathlib import Path
from typing import Any, Optional

import torch

from . import compat, config_adapter as ca
from .loader import DeviceBatchLoader
from .pipeline import DrawConfig, FeaturePlan


class TimedLoader:
    """Wraps a batch iterable and brackets the production of every batch with CUDA events on the current stream, so
    the feature stage's share of a train step can be read afterwards (``feature_ms()``)."""

    def __init__(self, loader):
        self.loader = loader
        self.events = []

    def __len__(self):
        return len(self.loader)

    def __iter__(self):
        it = iter(self.loader)
        while True:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            try:
                batch = next(it)
            except StopIteration:
                return
            e1.record()
            self.events.append((e0, e1))
            yield batch

    def feature_ms(self) -> float:
        """Sum of the device time of all batches produced so far (synchronises); clears the record."""
        torch.cuda.synchronize()
        ms = sum(a.elapsed_time(b) for a, b in self.events)
        self.events = []
        return ms


def synthetic_bank(n_clips: int, n_samples: int, device, seed: int = 0, pcm16: bool = True):
    """A device-resident clip bank of Gaussian 'speech' at 0.1 RMS with binary labels (no dataset ships with the task)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    wav = 0.1 * torch.randn(n_clips, n_samples, generator=g)
    labels = torch.randint(0, 2, (n_clips,), generator=g)
    if pcm16:
        wav = (wav.clamp(-1, 1) * 32767).to(torch.int16)
    return wav.to(device), labels.to(device)


def synthetic_aug_banks(n_noise: int, noise_len: int, n_rir: int, rir_len: int, seed: int = 1234):
    g = torch.Generator().manual_seed(seed)
    noise = [0.05 * torch.randn(noise_len, generator=g) for _ in range(n_noise)]
    t = torch.arange(rir_len, dtype=torch.float32)
    rirs = [torch.randn(rir_len, generator=g) * torch.exp(-t / 1000.0) for _ in range(n_rir)]
    return noise, rirs


def build_plan_and_loaders(config: Any, device, *, rank: int = 0, world_size: int = 1, n_train_clips: int = 1024,
                           n_val_clips: int = 0, n_samples: Optional[int] = None, seed: int = 0, noise=None, rirs=None):
    """WakewordConfig -> (FeaturePlan, train DeviceBatchLoader, val DeviceBatchLoader | []).  Feature and augmentation
    settings come from config.data / config.augmentation (config_adapter), batch size from config.training.batch_size
    PER RANK (src/config/defaults.py:35), SpecAugment with the reference test's parameters (2 + 2 masks, 15 / 35)."""
    n_samples = n_samples or ca.clip_samples(config)
    plan = ca.plan_from_config(config, device, spec_augment=True)
    if noise is None or rirs is None:
        noise, rirs = synthetic_aug_banks(64, n_samples, 16, 8000)
    plan.register_noise(noise)
    plan.register_rirs(rirs)
    draw = DrawConfig(**ca.draw_kwargs(config, seed=seed))
    bs = int(config.training.batch_size)
    bank, labels = synthetic_bank(n_train_clips, n_samples, plan.device, seed=seed + 1)
    train = DeviceBatchLoader(bank, labels, plan, bs, draw=draw, shuffle=True, seed=seed, rank=rank, world_size=world_size,
                              drop_last=True)
    val = []
    if n_val_clips > 0:
        vbank, vlabels = synthetic_bank(n_val_clips, n_samples, plan.device, seed=seed + 2)
        val = DeviceBatchLoader(vbank, vlabels, plan, bs, draw=None, shuffle=False, rank=rank, world_size=world_size)
    return plan, train, val


def build_ddp_trainer(config: Any, train_loader, val_loader, device, *, local_rank: int = 0, ddp: bool = True,
                      checkpoint_dir: Optional[str] = None):
    """The reference's model + Trainer around the given loaders.  With ``ddp`` the model is wrapped in
    DistributedDataParallel first (the default process group must exist); the Trainer itself is unmodified."""
    if not compat.reference_on_path():
        raise RuntimeError("the reference package (src.training, src.models) is not importable: neither /root/reference "
                           "nor baseline/_ref is present")
    compat.install_as_src_data()
    from src.models.architectures import create_model          # noqa: E402  (reference code)
    from src.training.trainer import Trainer                    # noqa: E402

    model = create_model(config.model.architecture, num_classes=config.model.num_classes, pretrained=False,
                         dropout=config.model.dropout, input_channels=1)
    model = model.to(device).to(memory_format=torch.channels_last)
    if ddp:
        from torch.nn.parallel import DistributedDataParallel as DDP
        model = DDP(model, device_ids=[local_rank], output_device=local_rank)
    ckpt = Path(checkpoint_dir or tempfile.mkdtemp(prefix="wwf_ddp_ckpt_"))
    return Trainer(model, train_loader, val_loader, config, checkpoint_dir=ckpt, device=str(device))


def grad_bytes(model: torch.nn.Module) -> int:
    return sum(p.numel() * p.element_size() for p in model.parameters() if p.requires_grad)


def time_allreduce(nbytes: int, device, reps: int = 10) -> float:
    """Milliseconds of one NCCL all-reduce of ``nbytes`` of float32 (the DDP gradient volume), CUDA events, after warm-up."""
    import torch.distributed as dist
    buf = torch.zeros(nbytes // 4, dtype=torch.float32, device=device)
    for _ in range(3):
        dist.all_reduce(buf)
    torch.cuda.synchronize(device)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        dist.all_reduce(buf)
    e1.record()
    torch.cuda.synchronize(device)
    return e0.elapsed_time(e1) / reps


def default_config(preset: str = "default"):
    """The reference's WakewordConfig (src/config/defaults.py:151) or one of its presets (src/config/presets.py:336-343)."""
    if not compat.reference_on_path():
        raise RuntimeError("the reference package is not importable")
    from src.config.defaults import WakewordConfig
    if preset in ("default", None):
        return WakewordConfig()
    from src.config.presets import PRESETS
    for name, make in PRESETS.items():
        if name.lower().startswith(preset.lower()):
            return make()
    raise KeyError(f"no preset named {preset!r}; have {list(PRESETS)}")


def _main():
    """torchrun entry: one DDP epoch on synthetic clips (used by tests/test_ddp_training.py and as a smoke run):
        python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 -m wakeword_trainer_home_b200.ddp_training"""
    import json
    import torch.distributed as dist
    rank, local_rank, world = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("LOCAL_RANK", 0), ("WORLD_SIZE", 1)))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29541")
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    cfg = default_config()
    cfg.training.batch_size = int(os.environ.get("WWF_DDP_BATCH", "32"))
    cfg.data.audio_duration = 1.5
    steps = int(os.environ.get("WWF_DDP_STEPS", "4"))
    plan, train, val = build_plan_and_loaders(cfg, dev, rank=rank, world_size=world,
                                              n_train_clips=steps * cfg.training.batch_size * world,
                                              n_val_clips=cfg.training.batch_size * world)
    trainer = build_ddp_trainer(cfg, train, val, dev, local_rank=local_rank)
    w0 = [p.detach().clone() for p in trainer.model.parameters()][:2]
    loss, acc = trainer.train_epoch(0)
    vloss, _ = trainer.validate_epoch(0)
    # after a DDP step every rank holds the same weights: compare a checksum across ranks
    chk = torch.stack([p.detach().double().sum() for p in trainer.model.parameters()]).sum().reshape(1)
    gathered = [torch.zeros_like(chk) for _ in range(world)]
    dist.all_gather(gathered, chk)
    moved = any(not torch.equal(a, b.detach()) for a, b in zip(w0, list(trainer.model.parameters())[:2]))
    if rank == 0:
        print(json.dumps({"ddp_epoch": {"world": world, "steps": steps, "train_loss": loss, "val_loss": vloss,
                                        "weights_equal_across_ranks": all(bool(torch.equal(g, gathered[0])) for g in gathered),
                                        "weights_moved": moved, "finite": bool(torch.isfinite(torch.tensor([loss, vloss])).all())}}),
              flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    _main()
