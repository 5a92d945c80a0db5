"""CPU tests of the oracle: both layers against the torchaudio golden vectors and each other,
plus the reference's own three assertions (tests/test_training_pipeline.py:242,243,262)."""
import os

import numpy as np
import pytest
import torch

from helpers import assert_features_close, aug_case_inputs, make_inputs
from oracle import np_oracle as npo
from oracle import ta_oracle as tao

CASES = [("feat_cfg1", 400, 40, 40), ("feat_refdefault", 1024, 128, 40), ("feat_n512", 512, 64, 32)]


@pytest.mark.parametrize("name,n_fft,n_mels,n_mfcc", CASES)
def test_ta_oracle_reproduces_golden(golden_dir, name, n_fft, n_mels, n_mfcc):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    x = make_inputs(int(g["seed"]), int(g["B"]), int(g["N"]))
    kw = dict(sample_rate=16000, n_mels=n_mels, n_mfcc=n_mfcc, n_fft=n_fft, hop_length=160)
    mel = tao.featurize(x, feature_type="mel", **kw).numpy()
    mf = tao.featurize(x, feature_type="mfcc", **kw).numpy()
    # same library, same machine class: float32 reductions may be re-associated by thread count
    assert np.abs(mel - g["logmel"]).max() <= 2e-4 and np.abs(mf - g["mfcc"]).max() <= 1e-3
    assert (mel[1] == -100.0).all()                      # silence -> exact clamp value


@pytest.mark.parametrize("name,n_fft,n_mels,n_mfcc", CASES)
def test_numpy_restatement_matches_torchaudio_f64(golden_dir, name, n_fft, n_mels, n_mfcc):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    x = make_inputs(int(g["seed"]), int(g["B"]), int(g["N"])).numpy()
    w = torch.hann_window(n_fft).numpy()
    kw = dict(sample_rate=16000, n_mels=n_mels, n_mfcc=n_mfcc, n_fft=n_fft, hop_length=160, fb=g["fb"], window=w)
    mel = npo.features(x, feature_type="mel", **kw)
    assert np.abs(mel - g["logmel64"]).max() <= 1e-9      # independent restatement == torchaudio in float64
    import torchaudio.functional as AF
    mf = npo.features(x, feature_type="mfcc", dct=AF.create_dct(n_mfcc, n_mels, "ortho").numpy(), **kw)
    for c in (0, 3):                                     # ordinary clips: within the path's tolerance of f32 torchaudio
        assert_features_close(mf[c], g["mfcc"][c], f"{name} mfcc clip {c}")


def test_numpy_filterbank_and_dct_constants():
    import torchaudio.functional as AF
    for n_freqs, n_mels in ((201, 40), (513, 128), (257, 64), (1025, 128), (129, 40)):
        fb = npo.mel_fbanks(n_freqs, 0.0, 8000.0, n_mels, 16000)
        ref = AF.melscale_fbanks(n_freqs, 0.0, 8000.0, n_mels, 16000).numpy()
        assert np.abs(fb - ref).max() <= 2e-5            # libm powf vs torch powf: last-ulp of f_pts, amplified
        assert ((fb > 0).sum(1) <= 2).all()              # at most two filters per FFT bin
    for n_mfcc, n_mels in ((40, 40), (40, 128), (13, 40)):
        assert np.abs(npo.create_dct(n_mfcc, n_mels) - AF.create_dct(n_mfcc, n_mels, "ortho").double().numpy()).max() <= 3e-6   # torchaudio's f32 cos
    assert np.abs(npo.hann_periodic(400) - torch.hann_window(400).double().numpy()).max() <= 5e-7  # torch computes it in float32


def test_augmentation_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "aug_cfg2.npz"))
    x, noise, rirs = aug_case_inputs(g)
    kw = dict(rirs=rirs, rir_idx=g["rir_idx"], noise_bank=noise, noise_idx=g["noise_idx"],
              noise_off=g["noise_off"], snr_db=torch.from_numpy(g["snr"]))
    y = tao.augment_wave(x, **kw).numpy()
    assert np.abs(y - g["mixed"]).max() <= 1e-5
    y64 = npo.add_noise(npo.rir_reverb(x.numpy(), [r.numpy() for r in rirs], g["rir_idx"]),
                        npo.gather_noise([n.numpy() for n in noise], g["noise_idx"], g["noise_off"], x.shape[1]),
                        g["snr"].astype(np.float64), active=g["noise_idx"] >= 0)
    rms = np.sqrt((g["mixed"].astype(np.float64) ** 2).mean(axis=1, keepdims=True))
    assert (np.abs(y64 - g["mixed"]) <= 2e-5 * rms).all()
    f = tao.pipeline(x, sample_rate=16000, feature_type="mfcc", n_mels=40, n_mfcc=40, n_fft=400, hop_length=160, **kw)
    assert_features_close(f.numpy(), g["mfcc"], "pipeline mfcc")
    dry = (g["rir_idx"] < 0) & (g["noise_idx"] < 0)
    assert np.array_equal(y[dry], x.numpy()[dry])


def test_masks_bit_exact_and_draws_in_range(golden_dir):
    g = np.load(os.path.join(golden_dir, "mask_ref.npz"))
    torch.manual_seed(int(g["seed"]))
    spec = torch.randn(4, 1, 64, 50)
    out = tao.spec_mask(spec, g["fstart"], g["flen"], g["tstart"], g["tlen"], 0.0)
    assert np.array_equal(out.numpy(), g["out"])
    assert np.array_equal(npo.spec_mask(spec.numpy(), g["fstart"], g["flen"], g["tstart"], g["tlen"], 0.0), g["out"])
    gen = torch.Generator().manual_seed(0)
    s, l = tao.draw_mask_params(gen, 1000, 64, 15, 2)
    assert (l >= 0).all() and (l < 15).all() and (s >= 0).all() and (s + l <= 64).all()


def test_draws_equal_torchaudio_mask_along_axis():
    """Same RNG stream -> the integer ranges equal what torchaudio's own mask_along_axis_iid masks."""
    import torchaudio.functional as AF
    spec = torch.ones(6, 1, 40, 151)
    torch.manual_seed(3)
    ref = AF.mask_along_axis_iid(spec, 35, 0.0, axis=3)
    torch.manual_seed(3)                                  # global generator, like torchaudio
    value = torch.rand(6, 1) * 35
    min_value = torch.rand(6, 1) * (151 - value)
    got = tao.spec_mask(spec, None, None, min_value.long().view(6, 1), value.long().view(6, 1), 0.0)
    assert torch.equal(got, ref)


def test_reference_assertions_on_reconstructed_classes():
    aug = tao.AudioAugmentation(sample_rate=16000, device="cpu", time_stretch_range=(0.8, 1.2),
                                pitch_shift_range=(-2, 2), background_noise_prob=0.5)
    t = torch.randn(1, 16000)
    a = aug(t)
    assert a.shape == t.shape and torch.isfinite(a).all()          # tests/test_training_pipeline.py:242-243
    s = torch.randn(1, 64, 50)
    assert tao.SpecAugment(15, 35, 2, 2)(s).shape == s.shape       # :262
    f = tao.FeatureExtractor(16000, "mel", 64, 40, 1024, 160)(torch.randn(16000))
    assert f.shape == (1, 64, 16000 // 160 + 1)                    # onnx_exporter.py:316-320


def test_top_db_is_per_clip():
    x = torch.stack([0.5 * torch.randn(8000), 1e-3 * torch.randn(8000)])
    both = tao.featurize(x, n_fft=400, n_mels=40, hop_length=160)
    alone = tao.featurize(x[1:], n_fft=400, n_mels=40, hop_length=160)
    assert torch.equal(both[1], alone[0])


# ---- time-stretch / pitch-shift / resample ------------------------------
def _shape_aug_inputs(g):
    N, B = int(g["N"]), int(g["B"])
    gen = torch.Generator().manual_seed(int(g["seed"]))
    x = 0.1 * torch.randn(B, N, generator=gen)
    t = torch.arange(N) / 16000.0
    x[1] = 0.5 * torch.sin(2 * torch.pi * 440.0 * t)
    return x


def _rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.linalg.norm(a - b) / np.linalg.norm(b)


def test_shape_aug_oracles_reproduce_torchaudio_golden(golden_dir):
    """Both oracle layers against vectors made by torchaudio's own _stretch_waveform / pitch_shift / resample.
    float64: tight.  float32: torchaudio's float32 phase cumsum makes its result reproducible only to its own
    float32-float64 gap (~1e-3), so the float32 comparison is bounded by that gap."""
    g = np.load(os.path.join(golden_dir, "shape_aug.npz"))
    x = _shape_aug_inputs(g)
    N = x.shape[1]
    w = torch.hann_window(512).numpy()       # the numpy layer is given torch's float32 window (see npo.stretch_core)
    for b, (n, rate, ln) in enumerate(zip(g["steps"], g["rates"], g["lens"])):
        n, rate, ln = int(n), float(rate), int(ln)
        xb = x[b:b + 1]
        s64 = tao.stretch_core(xb.double(), rate)[0].numpy()
        assert s64.shape[0] == ln and np.abs(s64 - g["stretch64"][b, :ln]).max() <= 1e-7
        assert np.abs(npo.stretch_core(xb.numpy(), rate, window=w)[0] - g["stretch64"][b, :ln]).max() <= 2e-6
        gap = _rel(g["stretch32"][b, :ln], g["stretch64"][b, :ln])
        assert _rel(tao.stretch_core(xb, rate)[0].numpy(), g["stretch32"][b, :ln]) <= 2 * gap
        assert _rel(npo.stretch_core(xb.numpy(), rate, np.float32, window=w)[0], g["stretch64"][b, :ln]) <= max(3 * gap, 5e-4)
        assert np.abs(tao.pitch_shift(xb.double(), [n])[0].numpy() - g["pitch64"][b]).max() <= 1e-7
        assert np.abs(npo.pitch_shift(xb.numpy(), [n], window=w)[0] - g["pitch64"][b]).max() <= 2e-6
        assert _rel(tao.pitch_shift(xb, [n])[0].numpy(), g["pitch32"][b]) <= 2 * _rel(g["pitch32"][b], g["pitch64"][b])
        # time_stretch = stretch_core cropped / zero-padded to N
        ts = tao.time_stretch(xb.double(), [rate])[0].numpy()
        assert np.array_equal(ts[:min(N, ln)], s64[:N]) and (ts[ln:] == 0).all()
    for o in (44100, 8000, 17959):
        want = g[f"rs_{o}"]
        assert np.abs(tao.resample(x[:2], o, 16000).numpy() - want).max() <= 1e-6
        # the numpy float32 layer follows torchaudio's float32 kernel arithmetic (sensitive for 17959:16000)
        assert np.abs(npo.resample(x[:2].numpy(), o, 16000, np.float32) - want).max() <= 2e-6
    assert torch.equal(tao.pitch_shift(x, [0] * x.shape[0]), x) and torch.equal(tao.time_stretch(x, [1.0] * x.shape[0]), x)
