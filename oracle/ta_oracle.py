"""torchaudio-backed oracle: the reference's (missing) ``src/data`` call surface,
reconstructed from its call sites and implemented with torchaudio's own CPU ops.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  This is also what
``bench.py --impl reference`` / ``cpu_baseline`` time: it is the reference's
"torchaudio CPU path" that BASELINE.json's north_star names.

Reconstruction evidence (paths relative to /root/reference):
* ``FeatureExtractor(sample_rate, feature_type, n_mels, n_mfcc, n_fft, hop_length, device)``
  and ``__call__(wave_1d) -> (1, F, T)``: src/evaluation/evaluator.py:86-94,125-128;
  src/evaluation/inference.py:94-102,197; T law src/export/onnx_exporter.py:316-320.
* ``AudioAugmentation(sample_rate, device, time_stretch_range, pitch_shift_range,
  background_noise_prob[, noise_snr_range, rir_prob])``: tests/test_training_pipeline.py:230-243;
  kwargs dict src/ui/panel_training.py:309-318; defaults src/config/defaults.py:73-95.
* ``SpecAugment(freq_mask_param, time_mask_param, n_freq_masks, n_time_masks)``:
  tests/test_training_pipeline.py:252-262.

Every random draw is hoisted out: the explicit-parameter functions (``featurize``,
``augment_wave``, ``spec_mask``, ``pipeline``) are what parity tests call; the classes
draw parameters with a torch.Generator and then call those.
"""
from __future__ import annotations

import torch
import torchaudio
import torchaudio.functional as AF
import torchaudio.transforms as AT

TA = "torchaudio " + torchaudio.__version__


def _transforms(sample_rate, feature_type, n_mels, n_mfcc, n_fft, hop_length, top_db, dtype):
    """MelSpectrogram + AmplitudeToDB('power', top_db) or MFCC(norm='ortho').

    TA/transforms/_transforms.py:566-632 (MelSpectrogram), :300-347 (AmplitudeToDB),
    :672-719 (MFCC; its internal AmplitudeToDB('power', 80.0) is :689-690)."""
    if feature_type in ("mel", "mel_spectrogram"):       # legacy alias, evaluator.py:82-83
        mel = AT.MelSpectrogram(sample_rate=sample_rate, n_fft=n_fft, hop_length=hop_length, n_mels=n_mels)
        db = AT.AmplitudeToDB("power", top_db=top_db)
        tr = torch.nn.Sequential(mel, db)
    elif feature_type == "mfcc":
        tr = AT.MFCC(sample_rate=sample_rate, n_mfcc=n_mfcc, norm="ortho", log_mels=False,
                     melkwargs=dict(n_fft=n_fft, hop_length=hop_length, n_mels=n_mels))
        tr.amplitude_to_DB.top_db = top_db
    else:
        raise ValueError(f"unknown feature_type {feature_type!r}")
    return tr.to(dtype)


def cmvn(feat: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    """Per-utterance CMVN (definition in oracle/np_oracle.py:cmvn; no torchaudio op)."""
    mu = feat.mean(dim=-1, keepdim=True)
    sd = (feat - mu).pow(2).mean(dim=-1, keepdim=True).sqrt()
    return (feat - mu) / (sd + eps)


@torch.no_grad()
def featurize(wave: torch.Tensor, *, sample_rate=16000, feature_type="mel", n_mels=128, n_mfcc=40,
              n_fft=1024, hop_length=160, top_db=80.0, use_cmvn=False, cmvn_eps=1e-5,
              dtype=torch.float32, _cache={}) -> torch.Tensor:
    """(B, N) -> (B, 1, F, T).  Fed to torchaudio as (B, 1, N) so that top_db is applied
    PER CLIP like the reference's per-sample calls (a (B, N) batch would share one
    cut-off, TA/functional/functional.py:393-402; SURVEY.md section 8c gotcha 1)."""
    key = (sample_rate, feature_type, n_mels, n_mfcc, n_fft, hop_length, top_db, dtype)
    if key not in _cache:
        _cache[key] = _transforms(*key)
    f = _cache[key](wave.to(dtype).unsqueeze(1))          # (B, 1, F, T)
    if use_cmvn:
        f = cmvn(f, cmvn_eps)
    return f


@torch.no_grad()
def rir_reverb(wave: torch.Tensor, rirs, rir_idx) -> torch.Tensor:
    """F.fftconvolve(x, h, 'full')[..., :N] per clip (TA/functional/functional.py:2222-2258).
    Clips sharing an RIR are convolved as one batched call."""
    out = wave.clone()
    N = wave.shape[-1]
    idx = torch.as_tensor(rir_idx)
    for r in idx.unique().tolist():
        if r < 0:
            continue
        sel = (idx == r).nonzero().flatten()
        h = torch.as_tensor(rirs[int(r)]).to(wave.dtype)
        out[sel] = AF.fftconvolve(wave[sel], h.unsqueeze(0), "full")[..., :N]
    return out


def gather_noise(bank, noise_idx, noise_off, N: int, dtype=torch.float32) -> torch.Tensor:
    """bank[idx][(off + j) mod len], j < N; zero rows where idx < 0 (index work, exact)."""
    idx = torch.as_tensor(noise_idx)
    out = torch.zeros(len(idx), N, dtype=dtype)
    ar = torch.arange(N)
    for b, i in enumerate(idx.tolist()):
        if i < 0:
            continue
        src = torch.as_tensor(bank[int(i)])
        out[b] = src[(int(noise_off[b]) + ar) % src.shape[0]].to(dtype)
    return out


@torch.no_grad()
def add_noise(wave, noise, snr_db, active=None) -> torch.Tensor:
    """F.add_noise(wave, noise, snr) (TA/functional/functional.py:2317-2382)."""
    y = AF.add_noise(wave, noise.to(wave.dtype), torch.as_tensor(snr_db).to(wave.dtype))
    if active is not None:
        y = torch.where(torch.as_tensor(active).unsqueeze(-1), y, wave)
    return y


# --------------------------------------------------------------------------------------
# A3: time-stretch / pitch-shift / resample
# --------------------------------------------------------------------------------------
PV_NFFT, PV_HOP = 512, 128      # F.pitch_shift defaults: n_fft=512, hop = n_fft // 4 (TA/functional/functional.py:1596-1604)


@torch.no_grad()
def stretch_core(wave: torch.Tensor, rate: float) -> torch.Tensor:
    """STFT -> F.phase_vocoder(rate) -> iSTFT(length=round(N / rate)): torchaudio's own
    ``_stretch_waveform`` (TA/functional/functional.py:1644-1693) with the rate given directly
    instead of through n_steps.  (B, N) -> (B, round(N / rate)), arithmetic in wave.dtype."""
    import math
    N = wave.shape[-1]
    # like _stretch_waveform, the window and phase_advance constants are built in float32 whatever wave.dtype is
    window = torch.hann_window(PV_NFFT).to(wave.dtype)
    spec = torch.stft(wave, PV_NFFT, PV_HOP, PV_NFFT, window, center=True, pad_mode="reflect",
                      normalized=False, onesided=True, return_complex=True)
    pa = torch.linspace(0, math.pi * PV_HOP, spec.shape[-2]).to(wave.dtype)[..., None]
    st = AF.phase_vocoder(spec, rate, pa)
    return torch.istft(st, PV_NFFT, PV_HOP, PV_NFFT, window, length=int(round(N / rate)))


def _fix_len(y: torch.Tensor, N: int) -> torch.Tensor:
    """_fix_waveform_shape (TA/functional/functional.py:1696-1718): crop or zero-pad to N."""
    return y[..., :N] if y.shape[-1] >= N else torch.nn.functional.pad(y, [0, N - y.shape[-1]])


@torch.no_grad()
def time_stretch(wave: torch.Tensor, rates) -> torch.Tensor:
    """Pitch-preserving speed change by rates[b] (> 1 = faster), shape kept: (B, N) -> (B, N).
    rate == 1.0 leaves the clip untouched.  Reference surface: ``time_stretch_range`` kwarg,
    tests/test_training_pipeline.py:233; AugmentationConfig.time_stretch_min/max, src/config/defaults.py:76-77."""
    out = wave.clone()
    for b, r in enumerate([float(v) for v in rates]):
        if r != 1.0:
            out[b] = _fix_len(stretch_core(wave[b:b + 1], r), wave.shape[-1])[0]
    return out


@torch.no_grad()
def pitch_shift(wave: torch.Tensor, n_steps, sample_rate: int = 16000) -> torch.Tensor:
    """F.pitch_shift(wave[b], sample_rate, n_steps[b]) per clip (TA/functional/functional.py:1596-1641);
    0 semitones leaves the clip untouched.  Reference surface: ``pitch_shift_range`` kwarg (integer
    semitones, src/config/validator.py:289-294), tests/test_training_pipeline.py:234."""
    out = wave.clone()
    steps = torch.as_tensor(n_steps)
    for n in steps.unique().tolist():
        if n == 0:
            continue
        sel = (steps == n).nonzero().flatten()
        out[sel] = AF.pitch_shift(wave[sel], sample_rate, int(n))
    return out


@torch.no_grad()
def resample(wave: torch.Tensor, orig_freq: int, new_freq: int) -> torch.Tensor:
    """F.resample, sinc_interp_hann, lowpass_filter_width 6, rolloff 0.99 (TA/functional/functional.py:1435-1497)."""
    return AF.resample(wave, orig_freq, new_freq)


@torch.no_grad()
def augment_wave(wave, *, rirs=None, rir_idx=None, noise_bank=None, noise_idx=None,
                 noise_off=None, snr_db=None, stretch_rate=None, pitch_steps=None, sample_rate=16000) -> torch.Tensor:
    """Time-domain half of the path: [time-stretch] -> [pitch-shift] -> RIR reverb -> noise @ SNR.
    (B, N) -> (B, N)."""
    x = wave
    if stretch_rate is not None:
        x = time_stretch(x, stretch_rate)
    if pitch_steps is not None:
        x = pitch_shift(x, pitch_steps, sample_rate)
    if rirs is not None and rir_idx is not None:
        x = rir_reverb(x, rirs, rir_idx)
    if noise_bank is not None and noise_idx is not None:
        nz = gather_noise(noise_bank, noise_idx, noise_off, x.shape[-1], x.dtype)
        x = add_noise(x, nz, snr_db, active=torch.as_tensor(noise_idx) >= 0)
    return x


def spec_mask(feat, fstart=None, flen=None, tstart=None, tlen=None, mask_value=0.0) -> torch.Tensor:
    """masked_fill of [start, start+len) along freq / time, the fill step of
    mask_along_axis(_iid) (TA/functional/functional.py:864-870,939-953).  Bit-exact."""
    out = feat.clone()
    F_, T_ = feat.shape[-2], feat.shape[-1]
    fa = torch.arange(F_).view(1, F_, 1)
    ta = torch.arange(T_).view(1, 1, T_)
    B = feat.shape[0]
    mask = torch.zeros(B, F_, T_, dtype=torch.bool)
    if fstart is not None:
        s = torch.as_tensor(fstart).view(B, -1); l = torch.as_tensor(flen).view(B, -1)
        for i in range(s.shape[1]):
            mask |= (fa >= s[:, i].view(B, 1, 1)) & (fa < (s[:, i] + l[:, i]).view(B, 1, 1))
    if tstart is not None:
        s = torch.as_tensor(tstart).view(B, -1); l = torch.as_tensor(tlen).view(B, -1)
        for i in range(s.shape[1]):
            mask |= (ta >= s[:, i].view(B, 1, 1)) & (ta < (s[:, i] + l[:, i]).view(B, 1, 1))
    while mask.dim() < out.dim():
        mask = mask.unsqueeze(1)
    return out.masked_fill(mask, mask_value)


@torch.no_grad()
def pipeline(wave, *, rirs=None, rir_idx=None, noise_bank=None, noise_idx=None, noise_off=None,
             snr_db=None, fstart=None, flen=None, tstart=None, tlen=None, mask_value=0.0,
             stretch_rate=None, pitch_steps=None, dtype=torch.float32, **feat_kw) -> torch.Tensor:
    """[stretch -> pitch ->] RIR -> noise -> features -> SpecAugment with every draw explicit."""
    x = augment_wave(wave.to(dtype), rirs=rirs, rir_idx=rir_idx, noise_bank=noise_bank,
                     noise_idx=noise_idx, noise_off=noise_off, snr_db=snr_db,
                     stretch_rate=stretch_rate, pitch_steps=pitch_steps,
                     sample_rate=feat_kw.get("sample_rate", 16000))
    f = featurize(x, dtype=dtype, **feat_kw)
    if fstart is not None or tstart is not None:
        f = spec_mask(f, fstart, flen, tstart, tlen, mask_value)
    return f


def draw_mask_params(gen: torch.Generator, B: int, size: int, mask_param: int, n_masks: int, p: float = 1.0):
    """The two uniform draws of mask_along_axis_iid turned into integers:
    value = U*param, min_value = U*(size - value); start = floor(min_value),
    len = floor(min_value)+floor(value) - start (TA/functional/functional.py:864-870)."""
    if p != 1.0:
        mask_param = min(mask_param, int(size * p))        # _get_mask_param, :806-810
    starts = torch.zeros(B, n_masks, dtype=torch.int32)
    lens = torch.zeros(B, n_masks, dtype=torch.int32)
    if mask_param < 1:
        return starts, lens
    for i in range(n_masks):
        value = torch.rand(B, generator=gen) * mask_param
        min_value = torch.rand(B, generator=gen) * (size - value)
        starts[:, i] = min_value.long().to(torch.int32)
        lens[:, i] = value.long().to(torch.int32)
    return starts, lens


# --------------------------------------------------------------------------------------
# Reconstructed classes
# --------------------------------------------------------------------------------------
class FeatureExtractor:
    def __init__(self, sample_rate=16000, feature_type="mel", n_mels=128, n_mfcc=40, n_fft=1024,
                 hop_length=160, device="cpu"):
        if feature_type == "mel_spectrogram":
            feature_type = "mel"
        self.kw = dict(sample_rate=sample_rate, feature_type=feature_type, n_mels=n_mels,
                       n_mfcc=n_mfcc, n_fft=n_fft, hop_length=hop_length)
        self.device = device

    def __call__(self, waveform: torch.Tensor) -> torch.Tensor:
        w = waveform.detach().to("cpu", torch.float32)
        single = w.dim() == 1 or (w.dim() == 2 and w.shape[0] == 1)
        w = w.reshape(-1, w.shape[-1])
        f = featurize(w, **self.kw)                        # (B, 1, F, T)
        return f[0] if single else f


class AudioAugmentation:
    def __init__(self, sample_rate=16000, device="cpu", time_stretch_range=(0.8, 1.2),
                 pitch_shift_range=(-2, 2), background_noise_prob=0.5, noise_snr_range=(5.0, 20.0),
                 rir_prob=0.25, background_noise=None, rirs=None, seed=0):
        self.sample_rate = sample_rate
        self.background_noise_prob = background_noise_prob
        self.noise_snr_range = noise_snr_range
        self.rir_prob = rir_prob
        self.noise_bank = background_noise
        self.rirs = rirs
        self.gen = torch.Generator().manual_seed(seed)

    def draw(self, B: int):
        g = self.gen
        rir_idx = torch.full((B,), -1, dtype=torch.int32)
        noise_idx = torch.full((B,), -1, dtype=torch.int32)
        noise_off = torch.zeros(B, dtype=torch.int64)
        lo, hi = self.noise_snr_range
        snr = lo + (hi - lo) * torch.rand(B, generator=g)
        if self.rirs:
            on = torch.rand(B, generator=g) < self.rir_prob
            pick = torch.randint(len(self.rirs), (B,), generator=g, dtype=torch.int32)
            rir_idx = torch.where(on, pick, rir_idx)
        if self.noise_bank:
            on = torch.rand(B, generator=g) < self.background_noise_prob
            pick = torch.randint(len(self.noise_bank), (B,), generator=g, dtype=torch.int32)
            noise_idx = torch.where(on, pick, noise_idx)
            lens = torch.tensor([len(self.noise_bank[i]) for i in pick.tolist()])
            noise_off = (torch.rand(B, generator=g) * lens).long()
        return dict(rir_idx=rir_idx, noise_idx=noise_idx, noise_off=noise_off, snr_db=snr)

    def __call__(self, waveform: torch.Tensor) -> torch.Tensor:
        shape = waveform.shape
        w = waveform.detach().to("cpu", torch.float32).reshape(-1, shape[-1])
        p = self.draw(w.shape[0])
        y = augment_wave(w, rirs=self.rirs, noise_bank=self.noise_bank, **p)
        return y.reshape(shape)


class SpecAugment:
    def __init__(self, freq_mask_param=15, time_mask_param=35, n_freq_masks=2, n_time_masks=2,
                 mask_value=0.0, seed=0):
        self.fp, self.tp, self.nf, self.nt = freq_mask_param, time_mask_param, n_freq_masks, n_time_masks
        self.mask_value = mask_value
        self.gen = torch.Generator().manual_seed(seed)

    def __call__(self, spec: torch.Tensor) -> torch.Tensor:
        s = spec.detach().to("cpu")
        lead = s.shape[:-2]
        s3 = s.reshape(-1, s.shape[-2], s.shape[-1])
        B, F_, T_ = s3.shape
        fs, fl = draw_mask_params(self.gen, B, F_, self.fp, self.nf)
        ts, tl = draw_mask_params(self.gen, B, T_, self.tp, self.nt)
        out = spec_mask(s3, fs, fl, ts, tl, self.mask_value)
        return out.reshape(*lead, F_, T_)
