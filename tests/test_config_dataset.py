"""The reference's own names on top of the path: config adapters (DataConfig / AugmentationConfig / presets),
``WakewordDataset`` / ``load_dataset_splits`` (src.data.dataset) and the ``src.data`` mount that lets the reference's
evaluator and trainer run unchanged.  CPU tests need no GPU; the -m gpu ones run the real kernels."""
import json
import os
import struct
import sys
import types

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from wakeword_trainer_home_b200 import compat, config_adapter as ca  # noqa: E402
from wakeword_trainer_home_b200.dataset import WakewordDataset, load_dataset_splits  # noqa: E402
from wakeword_trainer_home_b200.formats import save_split_manifest  # noqa: E402

HAVE_REF = compat.reference_on_path()      # /root/reference here, baseline/_ref on the GPU box


def _ns(**kw):
    return types.SimpleNamespace(**kw)


# the values of src/config/defaults.py:12-28,73-95, restated so that the adapter is tested even without the reference
DATA_DEFAULTS = dict(sample_rate=16000, audio_duration=2.5, n_mfcc=40, n_fft=1024, hop_length=160, n_mels=128,
                     feature_type="mel", normalize_audio=True)
AUG_DEFAULTS = dict(time_stretch_min=0.80, time_stretch_max=1.20, pitch_shift_min=-2, pitch_shift_max=2,
                    background_noise_prob=0.5, noise_snr_min=5.0, noise_snr_max=20.0, rir_prob=0.25,
                    freq_mask_prob=0.5, time_mask_prob=0.5)


def test_adapter_on_the_documented_defaults():
    cfg = _ns(data=_ns(**DATA_DEFAULTS), augmentation=_ns(**AUG_DEFAULTS))
    feat, aug, n = ca.summarize(cfg)
    assert feat == dict(sample_rate=16000, feature_type="mel", n_mels=128, n_mfcc=40, n_fft=1024, hop_length=160)
    # exactly the dict src/ui/panel_training.py:309-318 builds
    assert aug == dict(time_stretch_range=(0.8, 1.2), pitch_shift_range=(-2, 2), background_noise_prob=0.5,
                       noise_snr_range=(5.0, 20.0), rir_prob=0.25)
    assert n == 40000
    d = ca.draw_kwargs(cfg, seed=3)
    assert d["rir_prob"] == 0.25 and d["noise_prob"] == 0.5 and d["snr_range"] == (5.0, 20.0) and d["seed"] == 3
    assert d["freq_mask_prob"] == 0.5 and d["time_mask_prob"] == 0.5 and d["stretch_prob"] == 0.0
    # dicts and bare sections work too; 'mel_spectrogram' is the legacy alias (panel_training.py:321)
    assert ca.feature_kwargs(dict(DATA_DEFAULTS, feature_type="mel_spectrogram"))["feature_type"] == "mel"
    assert ca.feature_kwargs(_ns(**DATA_DEFAULTS)) == feat


def test_adapter_enforces_the_validator_envelope():
    with pytest.raises(ValueError, match=r"Hop length \(1024\) must be less than n_fft \(1024\)"):
        ca.feature_kwargs(dict(DATA_DEFAULTS, hop_length=1024))
    with pytest.raises(ValueError, match="Invalid feature type: spectrogram"):
        ca.feature_kwargs(dict(DATA_DEFAULTS, feature_type="spectrogram"))
    with pytest.raises(ValueError, match="n_fft=4096 is not built"):
        ca.feature_kwargs(dict(DATA_DEFAULTS, n_fft=4096))              # the validator allows it, the kernels do not
    assert any("Unusual FFT size: 400" in w for w in ca.validate_data_config(dict(DATA_DEFAULTS, n_fft=400)))
    with pytest.raises(ValueError, match="pitch_shift values must be integers"):
        ca.augmentation_kwargs(dict(AUG_DEFAULTS, pitch_shift_min=-1.5))
    with pytest.raises(ValueError, match=r"Probability must be in \[0, 1\]: 1.5"):
        ca.augmentation_kwargs(dict(AUG_DEFAULTS, rir_prob=1.5))
    with pytest.raises(ValueError, match="noise_snr"):
        ca.augmentation_kwargs(dict(AUG_DEFAULTS, noise_snr_min=20.0))
    with pytest.raises(ValueError, match="time_stretch"):
        ca.augmentation_kwargs(dict(AUG_DEFAULTS, time_stretch_min=1.3))
    assert ca.validate_augmentation_config(dict(AUG_DEFAULTS, time_stretch_min=0.4)) == ["Extreme time stretch range (0.5-2.0 recommended)"]


@pytest.mark.skipif(not HAVE_REF, reason="reference package not importable (neither /root/reference nor baseline/_ref)")
def test_every_reference_preset_round_trips():
    """The six presets of src/config/presets.py:336-343 and the bare defaults go through the adapters field by field."""
    from src.config.defaults import AugmentationConfig, DataConfig
    from src.config.presets import PRESETS
    assert ca.feature_kwargs(DataConfig())["n_fft"] == DataConfig().n_fft
    assert len(PRESETS) == 6
    for name, make in PRESETS.items():
        cfg = make()
        feat, aug, n = ca.summarize(cfg)
        d, a = cfg.data, cfg.augmentation
        assert feat == dict(sample_rate=d.sample_rate, feature_type=ca.normalize_feature_type(d.feature_type), n_mels=d.n_mels,
                            n_mfcc=d.n_mfcc, n_fft=d.n_fft, hop_length=d.hop_length), name
        assert aug == dict(time_stretch_range=(a.time_stretch_min, a.time_stretch_max),
                           pitch_shift_range=(a.pitch_shift_min, a.pitch_shift_max),
                           background_noise_prob=a.background_noise_prob,
                           noise_snr_range=(a.noise_snr_min, a.noise_snr_max), rir_prob=a.rir_prob), name
        assert n == int(d.sample_rate * d.audio_duration), name
        dk = ca.draw_kwargs(cfg)
        assert dk["noise_prob"] == a.background_noise_prob and dk["rir_prob"] == a.rir_prob
        assert dk["freq_mask_prob"] == a.freq_mask_prob and dk["time_mask_prob"] == a.time_mask_prob
    assert ca.augmentation_kwargs(AugmentationConfig())["rir_prob"] == 0.25


def _write_wav(path, x, rate):
    pcm = (np.clip(x, -1, 1) * 32767).astype("<i2")
    if pcm.ndim == 2:
        pcm = np.ascontiguousarray(pcm.T)
    ch = 1 if x.ndim == 1 else x.shape[0]
    data = pcm.tobytes()
    with open(path, "wb") as f:
        f.write(b"RIFF" + struct.pack("<I", 36 + len(data)) + b"WAVE" + b"fmt " +
                struct.pack("<IHHIIHH", 16, 1, ch, rate, rate * 2 * ch, 2 * ch, 16) + b"data" + struct.pack("<I", len(data)) + data)


def _make_tree(tmp_path, n=10, seed=0):
    rng = np.random.default_rng(seed)
    (tmp_path / "raw" / "positive").mkdir(parents=True)
    (tmp_path / "raw" / "negative").mkdir(parents=True)
    (tmp_path / "raw" / "background").mkdir(parents=True)
    (tmp_path / "raw" / "rirs").mkdir(parents=True)
    (tmp_path / "splits").mkdir()
    paths, labels = [], []
    for i in range(n):
        rate = (16000, 8000, 22050)[i % 3]
        dur = (0.6, 1.0, 1.4)[i % 3]                       # shorter and longer than the 1 s target
        x = 0.3 * rng.standard_normal((2, int(rate * dur))) if i % 4 == 0 else 0.3 * rng.standard_normal(int(rate * dur))
        lab = i % 2
        p = tmp_path / "raw" / ("positive" if lab else "negative") / f"clip_{i:02d}.wav"
        _write_wav(p, x, rate)
        paths.append(str(p.relative_to(tmp_path))); labels.append(lab)
    for j in range(3):
        _write_wav(tmp_path / "raw" / "background" / f"noise_{j}.wav", 0.05 * rng.standard_normal(20000), 16000)
    t = np.arange(2000)
    for j in range(2):
        _write_wav(tmp_path / "raw" / "rirs" / f"rir_{j}.wav", 0.5 * rng.standard_normal(2000) * np.exp(-t / 300.0), 16000)
    for name, sl in (("train", slice(0, n)), ("val", slice(0, n // 2)), ("test", slice(n // 2, n))):
        save_split_manifest(tmp_path / "splits" / f"{name}.json", paths[sl], labels[sl])
    return paths, labels


def test_dataset_reads_manifests_without_a_gpu(tmp_path):
    paths, labels = _make_tree(tmp_path)
    ds = WakewordDataset(tmp_path / "splits" / "train.json", sample_rate=16000, audio_duration=1.0, augment=False,
                         device="cuda", feature_type="mel_spectrogram", n_mels=40, n_mfcc=13, n_fft=400, hop_length=160,
                         data_root=tmp_path)
    assert len(ds) == len(paths) and ds.n_samples == 16000 and ds.feature_kw["feature_type"] == "mel"
    assert all(os.path.isabs(f) and os.path.exists(f) for f in ds.files)      # relative entries resolved against data_root
    assert ds.labels.tolist() == labels
    tr, va, te = load_dataset_splits(splits_dir=tmp_path / "splits", sample_rate=16000, audio_duration=1.0, augment_train=True,
                                     augmentation_config=ca.augmentation_kwargs(AUG_DEFAULTS), data_root=tmp_path, device="cuda",
                                     feature_type="mel", n_mels=40, n_mfcc=13, n_fft=400, hop_length=160)
    assert (len(tr), len(va), len(te)) == (10, 5, 5) and tr.augment and not va.augment and not te.augment
    assert tr.augmentation_config["rir_prob"] == 0.25
    os.remove(tmp_path / "splits" / "val.json")
    with pytest.raises(FileNotFoundError, match="val.json"):
        load_dataset_splits(tmp_path / "splits", data_root=tmp_path)


def test_src_data_mount_exposes_the_reference_names():
    compat.install_as_src_data(force=True)
    from src.data.augmentation import AudioAugmentation, SpecAugment
    from src.data.audio_utils import AudioProcessor
    from src.data.dataset import WakewordDataset as W2, load_dataset_splits as L2
    from src.data.feature_extraction import FeatureExtractor
    import wakeword_trainer_home_b200 as w
    assert FeatureExtractor is w.FeatureExtractor and AudioAugmentation is w.AudioAugmentation and SpecAugment is w.SpecAugment
    assert AudioProcessor is w.AudioProcessor and W2 is WakewordDataset and L2 is load_dataset_splits
    if HAVE_REF:
        import src.evaluation.evaluator as ev            # imports src.data.* at module level
        assert ev.FeatureExtractor is w.FeatureExtractor and ev.AudioProcessor is w.AudioProcessor


# ---------------------------------------------------------------------------------------------------------- GPU
@pytest.fixture(scope="module")
def ww():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import __graft_entry__ as ge
    ge.build()
    import wakeword_trainer_home_b200 as w
    return w


@pytest.mark.gpu
def test_dataset_items_equal_the_per_file_reference_chain(ww, tmp_path):
    """ds[i] = FeatureExtractor(AudioProcessor.process_audio(file)) - the chain the reference's evaluator runs per file
    (evaluator.py:119-128) - up to the int16 quantisation of the cached clip bank; item contract of evaluator.py:257-276."""
    _make_tree(tmp_path)
    kw = dict(sample_rate=16000, audio_duration=1.0, device="cuda", feature_type="mfcc", n_mels=40, n_mfcc=13, n_fft=400, hop_length=160)
    ds = WakewordDataset(tmp_path / "splits" / "train.json", augment=False, data_root=tmp_path, chunk=4, **kw)
    ap = ww.AudioProcessor(16000, 1.0, normalize=True)
    fe = ww.FeatureExtractor(16000, "mfcc", 40, 13, 400, 160, "cuda")
    for i in (0, 1, 5, 9, 3):
        f, lab, meta = ds[i]
        assert isinstance(lab, int) and lab == i % 2 and meta["path"].endswith(f"clip_{i:02d}.wav")
        assert f.device.type == "cpu" and f.dtype == torch.float32 and tuple(f.shape) == (1, 13, 101)
        wav = torch.from_numpy(ap.process_audio(ds.files[i]))
        pcm = (wav.clamp(-1, 32767 / 32768) * 32768).round() / 32768          # the bank keeps 16-bit PCM
        ref = fe(pcm).cpu()
        assert torch.allclose(f, ref, atol=2e-3, rtol=1e-4), (i, float((f - ref).abs().max()))
    # the reference's own DataLoader + collate (evaluator.py:257-277) works on it
    from torch.utils.data import DataLoader
    def collate(batch):
        feats, labels, metas = zip(*batch)
        return torch.stack(feats), torch.tensor(labels), list(metas)
    got = [b for b in DataLoader(ds, batch_size=4, shuffle=False, num_workers=0, pin_memory=True, collate_fn=collate)]
    assert [tuple(b[0].shape) for b in got] == [(4, 1, 13, 101), (4, 1, 13, 101), (2, 1, 13, 101)]
    # the batched loader yields the same features on the GPU
    fast = list(ds.loader(batch_size=4, shuffle=False))
    assert all(x.is_cuda for x, _ in fast)
    assert torch.allclose(torch.cat([x for x, _ in fast]).cpu(), torch.cat([b[0] for b in got]), atol=1e-5)
    assert torch.cat([y for _, y in fast]).cpu().tolist() == torch.cat([b[1] for b in got]).tolist()


@pytest.mark.gpu
def test_augmented_training_split_is_reproducible_and_finite(ww, tmp_path):
    _make_tree(tmp_path)
    cfg = dict(ca.augmentation_kwargs(dict(AUG_DEFAULTS, background_noise_prob=1.0, rir_prob=1.0)), freq_mask_prob=1.0, time_mask_prob=1.0)
    tr, va, te = load_dataset_splits(tmp_path / "splits", sample_rate=16000, audio_duration=1.0, augment_train=True,
                                     augmentation_config=cfg, data_root=tmp_path, device="cuda", feature_type="mel", n_mels=40,
                                     n_mfcc=13, n_fft=400, hop_length=160, spec_augment=True, chunk=10)
    a = torch.stack([tr[i][0] for i in range(len(tr))])
    assert tr.plan.n_noise == 3 and tr.plan.n_rir == 2                       # banks found under data_root/raw
    clean = torch.stack([WakewordDataset(tmp_path / "splits" / "train.json", 16000, 1.0, False, "cuda", "mel", 40, 13, 400, 160,
                                         data_root=tmp_path)[i][0] for i in range(len(tr))])
    assert torch.isfinite(a).all() and not torch.allclose(a, clean)
    assert (a == 0).any(), "SpecAugment masks (value 0) expected"
    tr2 = load_dataset_splits(tmp_path / "splits", sample_rate=16000, audio_duration=1.0, augment_train=True, augmentation_config=cfg,
                              data_root=tmp_path, device="cuda", feature_type="mel", n_mels=40, n_mfcc=13, n_fft=400, hop_length=160,
                              spec_augment=True, chunk=10)[0]
    assert torch.equal(a, torch.stack([tr2[i][0] for i in range(len(tr2))])), "same seed + epoch -> same draws"
    tr2.set_epoch(1)
    assert not torch.equal(a, torch.stack([tr2[i][0] for i in range(len(tr2))]))
    assert len(list(tr.loader(batch_size=4, shuffle=True, drop_last=True))) == 2


@pytest.mark.gpu
@pytest.mark.skipif(not HAVE_REF, reason="reference package not importable (baseline/_ref missing)")
def test_reference_evaluator_runs_unchanged_on_the_dataset(ww, tmp_path):
    """src/evaluation/evaluator.py:238-330 (ModelEvaluator.evaluate_dataset) and :100-160 (evaluate_file), unmodified,
    on top of the mounted src.data: AudioProcessor + FeatureExtractor + WakewordDataset are this package's."""
    compat.install_as_src_data(force=True)
    from src.evaluation.evaluator import ModelEvaluator
    from src.models.architectures import create_model
    _make_tree(tmp_path)
    torch.manual_seed(0)
    model = create_model("resnet18", num_classes=2, pretrained=False)
    kw = dict(sample_rate=16000, audio_duration=1.0, device="cuda", feature_type="mel", n_mels=64, n_mfcc=13, n_fft=512, hop_length=160)
    ev = ModelEvaluator(model, **kw)
    ds = WakewordDataset(tmp_path / "splits" / "test.json", augment=False, data_root=tmp_path, **kw)
    metrics, results = ev.evaluate_dataset(ds, threshold=0.5, batch_size=4)
    assert len(results) == len(ds) == 5
    assert {r.filename for r in results} == {os.path.basename(p) for p in ds.paths}
    assert all(np.isfinite(r.confidence) and r.prediction in ("Positive", "Negative") for r in results)
    assert 0.0 <= metrics.accuracy <= 1.0
    from pathlib import Path
    one = ev.evaluate_file(Path(ds.files[0]))
    assert one.filename == os.path.basename(ds.files[0]) and np.isfinite(one.confidence)
