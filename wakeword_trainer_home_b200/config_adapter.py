"""Build the path's objects from the reference's own configuration dataclasses.

The reference keeps every knob of this path in ``DataConfig`` (src/config/defaults.py:12-28: sample_rate,
audio_duration, n_mfcc, n_fft, hop_length, n_mels, feature_type, normalize_audio) and ``AugmentationConfig``
(src/config/defaults.py:73-95: time_stretch_min/max, pitch_shift_min/max, background_noise_prob, noise_snr_min/max,
rir_prob, freq_mask_prob, time_mask_prob); its six presets (src/config/presets.py:336-343) are instances of the
same dataclasses.  The adapters below accept those objects (or any object / dict with the same field names) verbatim,
apply the envelope the reference's validator enforces for them (src/config/validator.py:129-151,266-314: the
*errors* raise here, the *warnings* are returned) and produce ``FeaturePlan`` / ``DrawConfig`` /
``AudioAugmentation`` keyword arguments.  Nothing in here needs a GPU.
"""
from __future__ import annotations

from typing import Any, Dict, List, Optional, Tuple

from . import _native as N

VALID_FEATURES = ("mel", "mel_spectrogram", "mfcc")           # validator.py:145 ('mel_spectrogram' is the legacy alias)
TYPICAL_FFT_SIZES = (256, 512, 1024, 2048, 4096)              # validator.py:129


def _get(cfg: Any, name: str, default=None):
    if isinstance(cfg, dict):
        return cfg.get(name, default)
    return getattr(cfg, name, default)


def _section(cfg: Any, name: str):
    """``cfg`` may be a whole WakewordConfig (has .data / .augmentation) or the section itself."""
    sub = _get(cfg, name)
    return sub if sub is not None else cfg


def normalize_feature_type(feature_type: str) -> str:
    """'mel_spectrogram' -> 'mel' (src/ui/panel_training.py:321, src/evaluation/evaluator.py:82-83)."""
    return "mel" if feature_type == "mel_spectrogram" else feature_type


def validate_data_config(data_cfg: Any) -> List[str]:
    """The validator's rules for the fields this path consumes.  Errors raise ValueError with the validator's wording
    (validator.py:138-151); its warnings (n_fft outside the typical set, n_mfcc range) come back as strings.  On top,
    the envelope of the CUDA kernels: n_fft in {256, 400, 512, 1024, 2048}, 1 <= n_mels <= 128."""
    data = _section(data_cfg, "data")
    n_fft, hop = int(_get(data, "n_fft", 1024)), int(_get(data, "hop_length", 160))
    ft = _get(data, "feature_type", "mel")
    warnings = []
    if hop >= n_fft:
        raise ValueError(f"Hop length ({hop}) must be less than n_fft ({n_fft})")
    if ft not in VALID_FEATURES:
        raise ValueError(f"Invalid feature type: {ft} (valid: {list(VALID_FEATURES)})")
    if n_fft not in N.SUPPORTED_N_FFT:
        raise ValueError(f"n_fft={n_fft} is not built for the B200 path (supported: {list(N.SUPPORTED_N_FFT)})")
    if n_fft not in TYPICAL_FFT_SIZES:
        warnings.append(f"Unusual FFT size: {n_fft} (typical: {list(TYPICAL_FFT_SIZES)})")
    n_mfcc = int(_get(data, "n_mfcc", 40))
    if n_mfcc < 13:
        warnings.append(f"Low MFCC count: {n_mfcc} (13-40 recommended)")
    n_mels = int(_get(data, "n_mels", 128))
    if not 1 <= n_mels <= 128:
        raise ValueError(f"n_mels={n_mels} must be in [1, 128] for the B200 path")
    if normalize_feature_type(ft) == "mfcc" and n_mfcc > n_mels:
        raise ValueError(f"n_mfcc={n_mfcc} must not exceed n_mels={n_mels}")
    return warnings


def validate_augmentation_config(aug_cfg: Any) -> List[str]:
    """validator.py:266-314: min < max for stretch / pitch / SNR, integer semitones, probabilities in [0, 1]."""
    a = _section(aug_cfg, "augmentation")
    warnings = []
    ts = (float(_get(a, "time_stretch_min", 0.8)), float(_get(a, "time_stretch_max", 1.2)))
    if ts[0] >= ts[1]:
        raise ValueError(f"augmentation.time_stretch: min ({ts[0]}) must be less than max ({ts[1]})")
    if ts[0] < 0.5 or ts[1] > 2.0:
        warnings.append("Extreme time stretch range (0.5-2.0 recommended)")
    ps = (_get(a, "pitch_shift_min", -2), _get(a, "pitch_shift_max", 2))
    if not isinstance(ps[0], int) or not isinstance(ps[1], int):
        raise ValueError(f"augmentation.pitch_shift: pitch_shift values must be integers (got min={type(ps[0]).__name__}, "
                         f"max={type(ps[1]).__name__})")
    if ps[0] >= ps[1]:
        raise ValueError(f"augmentation.pitch_shift: min ({ps[0]}) must be less than max ({ps[1]})")
    for field in ("background_noise_prob", "rir_prob", "freq_mask_prob", "time_mask_prob"):
        v = float(_get(a, field, 0.0))
        if not 0 <= v <= 1:
            raise ValueError(f"augmentation.{field}: Probability must be in [0, 1]: {v}")
    snr = (float(_get(a, "noise_snr_min", 5.0)), float(_get(a, "noise_snr_max", 20.0)))
    if snr[0] >= snr[1]:
        raise ValueError(f"augmentation.noise_snr: min ({snr[0]}) must be less than max ({snr[1]})")
    return warnings


def feature_kwargs(data_cfg: Any) -> Dict[str, Any]:
    """DataConfig -> the FeatureExtractor / FeaturePlan constructor arguments (evaluator.py:86-94)."""
    validate_data_config(data_cfg)
    data = _section(data_cfg, "data")
    return dict(sample_rate=int(_get(data, "sample_rate", 16000)),
                feature_type=normalize_feature_type(_get(data, "feature_type", "mel")),
                n_mels=int(_get(data, "n_mels", 128)), n_mfcc=int(_get(data, "n_mfcc", 40)),
                n_fft=int(_get(data, "n_fft", 1024)), hop_length=int(_get(data, "hop_length", 160)))


def clip_samples(data_cfg: Any) -> int:
    """Samples per clip: sample_rate * audio_duration (evaluator.py:76-79,119-122)."""
    data = _section(data_cfg, "data")
    return int(int(_get(data, "sample_rate", 16000)) * float(_get(data, "audio_duration", 2.5)))


def augmentation_kwargs(aug_cfg: Any) -> Dict[str, Any]:
    """AugmentationConfig -> the kwargs dict the reference hands to its loaders (src/ui/panel_training.py:309-318)."""
    validate_augmentation_config(aug_cfg)
    a = _section(aug_cfg, "augmentation")
    return dict(time_stretch_range=(float(_get(a, "time_stretch_min", 0.8)), float(_get(a, "time_stretch_max", 1.2))),
                pitch_shift_range=(int(_get(a, "pitch_shift_min", -2)), int(_get(a, "pitch_shift_max", 2))),
                background_noise_prob=float(_get(a, "background_noise_prob", 0.5)),
                noise_snr_range=(float(_get(a, "noise_snr_min", 5.0)), float(_get(a, "noise_snr_max", 20.0))),
                rir_prob=float(_get(a, "rir_prob", 0.25)))


def draw_kwargs(aug_cfg: Any, *, seed: int = 0, shape_aug_prob: float = 0.0, freq_mask_param: int = 15,
                time_mask_param: int = 35) -> Dict[str, Any]:
    """AugmentationConfig -> ``DrawConfig`` fields for the on-GPU draws.  ``freq_mask_prob`` / ``time_mask_prob`` gate
    the whole SpecAugment op per clip; mask sizes are torchaudio / the reference test's
    (tests/test_training_pipeline.py:252-257).  ``shape_aug_prob``: apply probability of time-stretch and pitch-shift
    (the reference's is unknowable, its module is absent; 0 keeps them off)."""
    validate_augmentation_config(aug_cfg)
    a = _section(aug_cfg, "augmentation")
    return dict(seed=seed, rir_prob=float(_get(a, "rir_prob", 0.25)), noise_prob=float(_get(a, "background_noise_prob", 0.5)),
                freq_mask_prob=float(_get(a, "freq_mask_prob", 0.5)), time_mask_prob=float(_get(a, "time_mask_prob", 0.5)),
                snr_range=(float(_get(a, "noise_snr_min", 5.0)), float(_get(a, "noise_snr_max", 20.0))),
                freq_mask_param=int(freq_mask_param), time_mask_param=int(time_mask_param),
                stretch_prob=float(shape_aug_prob),
                stretch_range=(float(_get(a, "time_stretch_min", 0.8)), float(_get(a, "time_stretch_max", 1.2))),
                pitch_prob=float(shape_aug_prob),
                pitch_range=(int(_get(a, "pitch_shift_min", -2)), int(_get(a, "pitch_shift_max", 2))))


def plan_from_config(config: Any, device="cuda", *, spec_augment: bool = False, out_dtype=None, **overrides):
    """WakewordConfig / DataConfig -> FeaturePlan (needs the CUDA library)."""
    from .pipeline import FeaturePlan
    import torch
    kw = feature_kwargs(config)
    if spec_augment:
        kw.update(n_freq_masks=2, n_time_masks=2)
    if out_dtype is not None:
        kw["out_dtype"] = out_dtype
    kw.update(overrides)
    return FeaturePlan(device=device, **kw) if "out_dtype" in kw else FeaturePlan(device=device, out_dtype=torch.float32, **kw)


def draw_config_from(config: Any, **kw):
    from .pipeline import DrawConfig
    return DrawConfig(**draw_kwargs(config, **kw))


def summarize(config: Any) -> Tuple[Dict[str, Any], Dict[str, Any], int]:
    """(feature kwargs, augmentation kwargs, samples per clip) of a WakewordConfig - what a preset means for this path."""
    return feature_kwargs(config), augmentation_kwargs(config), clip_samples(config)
