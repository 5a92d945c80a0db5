"""Loader resume state (state_dict / load_state_dict), the (B, T, F) sequence view and the NVTX switch.
The checkpoint dict of the reference's trainer (src/training/trainer.py:486-525) carries model / optimizer / scheduler
state; the loaders' state is what has to sit beside it for a resumed run to see identical batches."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from helpers import synth_banks  # noqa: E402


def test_as_sequence_is_the_lstm_view():
    import wakeword_trainer_home_b200 as ww
    x = torch.arange(2 * 1 * 3 * 5, dtype=torch.float32).reshape(2, 1, 3, 5)
    s = ww.as_sequence(x)
    assert s.shape == (2, 5, 3) and s.data_ptr() == x.data_ptr()          # a view: (batch, time_steps, features)
    assert torch.equal(s[1, 4], x[1, 0, :, 4])
    with pytest.raises(ValueError):
        ww.as_sequence(torch.zeros(2, 3, 5))


@pytest.fixture(scope="module")
def ww():
    import wakeword_trainer_home_b200 as w
    return w


@pytest.mark.gpu
def test_device_loader_resumes_mid_epoch_with_identical_batches(ww):
    gen = torch.Generator().manual_seed(11)
    n, N, B = 96, 16000, 16
    bank = (0.1 * torch.randn(n, N, generator=gen)).cuda()
    labels = torch.randint(0, 2, (n,), generator=gen)
    noise, rirs = synth_banks(3, 4, 20000, 3, 3000)

    def make():
        plan = ww.FeaturePlan(16000, "mel", 40, 40, 400, 160, "cuda", n_freq_masks=2, n_time_masks=2)
        plan.register_noise(noise); plan.register_rirs(rirs)
        return ww.DeviceBatchLoader(bank, labels, plan, B, draw=ww.DrawConfig(seed=7), shuffle=True, seed=5)

    full = make()
    full.set_epoch(2)
    ref = [(f.clone(), t.clone()) for f, t in full]
    assert len(ref) == 6
    # run 4 batches, checkpoint, restore into a fresh loader, continue
    a = make()
    a.set_epoch(2)
    it = iter(a)
    for _ in range(4):
        next(it)
    state = a.state_dict()
    assert state["batches_done"] == 4 and state["samples_drawn"] == 4 * B and state["epoch"] == 2
    b = make()
    b.load_state_dict(state)
    rest = [(f.clone(), t.clone()) for f, t in b]
    assert len(rest) == 2
    for (f0, t0), (f1, t1) in zip(ref[4:], rest):
        assert torch.equal(f0, f1) and torch.equal(t0, t1)
    # the next epoch of the restored loader equals the next epoch of the uninterrupted one
    full.set_epoch(3); b.set_epoch(3)
    for (f0, t0), (f1, t1) in zip(full, b):
        assert torch.equal(f0, f1) and torch.equal(t0, t1)
    with pytest.raises(ValueError):
        b.load_state_dict({"kind": "GpuBatchLoader"})


@pytest.mark.gpu
def test_host_loader_state_and_nvtx_switch(ww):
    gen = torch.Generator().manual_seed(12)
    n, N, B = 40, 16000, 8
    clips = 0.1 * torch.randn(n, N, generator=gen)
    labels = torch.randint(0, 2, (n,), generator=gen)
    noise, rirs = synth_banks(4, 3, 20000, 2, 3000)
    plan = ww.FeaturePlan(16000, "mfcc", 40, 13, 400, 160, "cuda", n_freq_masks=2, n_time_masks=2)
    plan.nvtx(True)                                                       # ranges on: results unchanged
    aug = ww.AudioAugmentation(16000, "cuda", background_noise_prob=0.5, rir_prob=0.5, background_noise=noise, rirs=rirs,
                               plan=plan, seed=1)
    sa = ww.SpecAugment(15, 35, 2, 2)
    mk = lambda: ww.GpuBatchLoader(clips, labels, plan, B, augment=aug, spec_augment=sa, shuffle=True, seed=9)
    full = mk()
    ref = [f.clone() for f, _ in full]
    a = mk()
    it = iter(a)
    for _ in range(2):
        next(it)
    st = a.state_dict()
    assert st["step"] == 2 and st["batches_done"] == 2
    b = mk()
    b.load_state_dict(st)
    got = [f.clone() for f, _ in b]
    assert len(got) == len(ref) - 2
    for f0, f1 in zip(ref[2:], got):
        assert torch.equal(f0, f1)
    plan.nvtx(False)
    x = clips[:4].cuda()
    assert torch.equal(plan.featurize(x), plan.featurize(x))
    assert ww.as_sequence(plan.featurize(x)).shape == (4, N // 160 + 1, 13)
