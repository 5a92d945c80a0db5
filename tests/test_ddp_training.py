"""BASELINE.json configs[2]: DeviceBatchLoader -> the reference's own Trainer.train_epoch + create_model('resnet18')
under DistributedDataParallel / NCCL (wakeword_trainer_home_b200/ddp_training.py).  The reference package is imported
unmodified from /root/reference (build container) or baseline/_ref (GPU box)."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from wakeword_trainer_home_b200 import compat  # noqa: E402

HAVE_REF = compat.reference_on_path()


def _run_ddp(nproc: int, port: int, steps: int = 3):
    env = dict(os.environ, WWF_DDP_STEPS=str(steps), WWF_DDP_BATCH="16", MASTER_PORT=str(port), PYTHONPATH=ROOT)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), "-m", "wakeword_trainer_home_b200.ddp_training"]
    r = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith('{"ddp_epoch"')]
    assert lines, r.stdout[-2000:] + r.stderr[-2000:]
    return json.loads(lines[-1])["ddp_epoch"]


@pytest.mark.skipif(not HAVE_REF, reason="reference package not importable")
def test_reference_config_and_model_import_without_gpu():
    """The harness builds the reference's config and classifier from the reference's own modules (CPU part only)."""
    from wakeword_trainer_home_b200 import ddp_training as dt
    cfg = dt.default_config()
    assert cfg.model.architecture == "resnet18" and cfg.training.batch_size == 128 and cfg.data.n_fft == 1024
    edge = dt.default_config("edge")
    assert edge.model.architecture == "mobilenetv3" and edge.data.audio_duration == 1.2      # presets.py (BASELINE.json's "2 s, 64 mels" is not what ships: SURVEY App. C)
    compat.install_as_src_data()
    from src.models.architectures import create_model
    m = create_model("resnet18", num_classes=2, pretrained=False, input_channels=1)
    assert dt.grad_bytes(m) == 4 * sum(p.numel() for p in m.parameters())
    assert 44e6 < dt.grad_bytes(m) < 45.5e6           # the 44.7 MB fp32 gradient all-reduce of configs[2]


@pytest.mark.gpu
@pytest.mark.skipif(not HAVE_REF, reason="reference package not importable")
def test_one_ddp_epoch_single_rank():
    """World size 1 still goes through init_process_group('nccl') + DistributedDataParallel + the reference Trainer."""
    r = _run_ddp(1, 29611)
    assert r["finite"] and r["weights_moved"] and r["world"] == 1, r


@pytest.mark.gpu
@pytest.mark.skipif(not HAVE_REF, reason="reference package not importable")
@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
def test_one_ddp_epoch_two_ranks():
    """Two ranks, each with its own clip shard and its own feature plan; after the epoch both hold identical weights
    (the gradient all-reduce ran) and the loss is finite."""
    r = _run_ddp(2, 29612)
    assert r["finite"] and r["weights_moved"] and r["weights_equal_across_ranks"] and r["world"] == 2, r
