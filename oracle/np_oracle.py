"""numpy restatement of the torchaudio arithmetic behind the reference's feature path.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Every function cites the
torchaudio source it follows (``TA/`` = site-packages/torchaudio, v2.11.0 here; the
reference pins 2.1.2, ``/root/reference/requirements.txt:6``) and the reference call
site that makes it part of the path.

All functions take/return numpy arrays.  ``dtype`` selects the arithmetic type:
float64 (default; the "true value" both the CUDA path and torchaudio-fp32 are
compared against) or float32.
"""
from __future__ import annotations

import math
import numpy as np

__all__ = [
    "hann_periodic", "stft_power", "mel_fbanks", "create_dct", "amplitude_to_db",
    "log_mel", "mfcc", "cmvn", "rir_reverb", "gather_noise", "add_noise",
    "spec_mask", "num_frames", "features", "pipeline",
    "stft_complex", "phase_vocoder", "istft", "stretch_core", "time_stretch", "sinc_resample_kernel",
    "resample", "pitch_shift",
]


def num_frames(n_samples: int, hop_length: int) -> int:
    """Frame count law stated by the reference at src/export/onnx_exporter.py:316-317."""
    return n_samples // hop_length + 1


def hann_periodic(n_fft: int, dtype=np.float64) -> np.ndarray:
    """torch.hann_window(n_fft, periodic=True): w[n] = 0.5 - 0.5 cos(2 pi n / n_fft).

    Window default of Spectrogram, TA/transforms/_transforms.py:64-77."""
    n = np.arange(n_fft, dtype=np.float64)
    return (0.5 - 0.5 * np.cos(2.0 * np.pi * n / n_fft)).astype(dtype)


def stft_power(x: np.ndarray, n_fft: int, hop_length: int, dtype=np.float64, window=None) -> np.ndarray:
    """|STFT|^2 with center=True, reflect pad, periodic Hann, onesided.

    Follows F.spectrogram -> torch.stft, TA/functional/functional.py:123-145
    (power=2.0, normalized=False, win_length=n_fft).  x: (B, N) -> (B, K, T).
    ``window`` may inject torch.hann_window's float32 values: torchaudio keeps the window as
    a float32 buffer even in a .double() module, and on a pure tone its rounding moves the
    -78 dB side lobes by ~1e-3 dB."""
    x = np.asarray(x, dtype=dtype)
    B, N = x.shape
    pad = n_fft // 2
    if N <= pad:
        raise ValueError("reflect padding needs N > n_fft/2")
    xp = np.pad(x, ((0, 0), (pad, pad)), mode="reflect")
    T = num_frames(N, hop_length)
    idx = (np.arange(T) * hop_length)[:, None] + np.arange(n_fft)[None, :]
    w = hann_periodic(n_fft, dtype) if window is None else np.asarray(window).astype(dtype)
    frames = xp[:, idx] * w[None, None, :]                                 # (B, T, n_fft)
    spec = np.fft.rfft(frames.astype(np.float64), axis=-1)                 # numpy FFT is f64
    if dtype == np.float32:
        spec = spec.astype(np.complex64)
    power = (spec.real ** 2 + spec.imag ** 2).astype(dtype)
    return np.transpose(power, (0, 2, 1))                                  # (B, K, T)


def _linspace32(start: float, end: float, steps: int) -> np.ndarray:
    """torch.linspace in float32: forward from start for the first half, backward from
    end for the second half (ATen RangeFactories linspace kernel)."""
    start = np.float32(start)
    end = np.float32(end)
    step = np.float32((end - start) / np.float32(steps - 1))
    i = np.arange(steps)
    half = steps // 2
    lo = (start + step * i.astype(np.float32)).astype(np.float32)
    hi = (end - step * (steps - 1 - i).astype(np.float32)).astype(np.float32)
    return np.where(i < half, lo, hi).astype(np.float32)


def mel_fbanks(n_freqs: int, f_min: float, f_max: float, n_mels: int, sample_rate: int) -> np.ndarray:
    """HTK mel triangular filterbank, norm=None -> (n_freqs, n_mels) float32.

    Restates melscale_fbanks / _create_triangular_filterbank / _hz_to_mel / _mel_to_hz,
    TA/functional/functional.py:425-447,450-487,490-513,518-587, in float32 with the
    same operation order (it is a float32 constant in torchaudio)."""
    all_freqs = _linspace32(0.0, float(sample_rate // 2), n_freqs)
    m_min = 2595.0 * math.log10(1.0 + (f_min / 700.0))
    m_max = 2595.0 * math.log10(1.0 + (f_max / 700.0))
    m_pts = _linspace32(m_min, m_max, n_mels + 2)
    f_pts = (np.float32(700.0) * (np.power(np.float32(10.0), m_pts / np.float32(2595.0)) - np.float32(1.0))).astype(np.float32)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts[None, :] - all_freqs[:, None]
    down = (np.float32(-1.0) * slopes[:, :-2]) / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    return np.maximum(np.float32(0.0), np.minimum(down, up)).astype(np.float32)


def create_dct(n_mfcc: int, n_mels: int, dtype=np.float64) -> np.ndarray:
    """Orthonormal DCT-II matrix (n_mels, n_mfcc): TA/functional/functional.py:636-667."""
    n = np.arange(n_mels, dtype=np.float64)
    k = np.arange(n_mfcc, dtype=np.float64)[:, None]
    dct = np.cos(math.pi / float(n_mels) * (n + 0.5) * k)
    dct[0] *= 1.0 / math.sqrt(2.0)
    dct *= math.sqrt(2.0 / float(n_mels))
    return dct.T.astype(dtype)


def amplitude_to_db(x: np.ndarray, top_db: float | None = 80.0) -> np.ndarray:
    """AmplitudeToDB('power'): 10 log10(clamp(x, 1e-10)), then a PER-CLIP floor at
    max - top_db.  TA/functional/functional.py:390-404 (multiplier 10, amin 1e-10,
    ref 1.0 -> db_multiplier 0).  x: (B, F, T); the cut-off is per leading index, which
    is what the reference's per-sample calls produce (SURVEY.md section 8c gotcha 1)."""
    x_db = 10.0 * np.log10(np.maximum(x, x.dtype.type(1e-10)))
    if top_db is not None and top_db >= 0:
        cut = x_db.max(axis=(-2, -1), keepdims=True) - x_db.dtype.type(top_db)
        x_db = np.maximum(x_db, cut)
    return x_db


def log_mel(x, sample_rate, n_fft, hop_length, n_mels, top_db=80.0, f_min=0.0, f_max=None, dtype=np.float64,
            fb=None, window=None):
    """MelSpectrogram + AmplitudeToDB.  mel = (spec^T @ fb)^T, TA/transforms/_transforms.py:417.
    ``fb`` may inject torchaudio's own float32 filterbank (its last ulp depends on the powf
    implementation, and the triangle slopes amplify that to ~1e-5 relative)."""
    f_max = float(sample_rate // 2) if f_max is None else f_max
    power = stft_power(x, n_fft, hop_length, dtype, window)                # (B, K, T)
    if fb is None:
        fb = mel_fbanks(n_fft // 2 + 1, f_min, f_max, n_mels, sample_rate)
    fb = np.asarray(fb).astype(dtype)
    mel = np.einsum("bkt,km->bmt", power, fb).astype(dtype)
    return amplitude_to_db(mel, top_db)


def mfcc(x, sample_rate, n_fft, hop_length, n_mels, n_mfcc, top_db=80.0, dtype=np.float64, dct=None, **kw):
    """MFCC(log_mels=False, norm='ortho'): DCT-II of the dB mel, TA/transforms/_transforms.py:672-719.
    ``dct`` may inject torchaudio's float32 matrix: its cos() of arguments up to ~120 rad is only
    good to ~1.3e-6, which times |dB| ~ 100 over n_mels terms is visible at the 1e-3 level."""
    mel_db = log_mel(x, sample_rate, n_fft, hop_length, n_mels, top_db, dtype=dtype, **kw)
    dct = create_dct(n_mfcc, n_mels, dtype) if dct is None else np.asarray(dct).astype(dtype)
    return np.einsum("bmt,mc->bct", mel_db, dct).astype(dtype)


def cmvn(feat: np.ndarray, eps: float = 1e-5) -> np.ndarray:
    """Per-utterance CMVN.  Named only by BASELINE.json north_star; there is no
    reference call site or torchaudio op for it, so this is
    the definition both sides implement: per clip and per coefficient over time,
    (x - mean) / (population_std + eps)."""
    mu = feat.mean(axis=-1, keepdims=True)
    sd = np.sqrt(((feat - mu) ** 2).mean(axis=-1, keepdims=True))
    return (feat - mu) / (sd + feat.dtype.type(eps))


def rir_reverb(x: np.ndarray, rirs: list[np.ndarray], rir_idx: np.ndarray, dtype=np.float64) -> np.ndarray:
    """Full linear convolution with the selected RIR, cropped to the first N samples.

    F.fftconvolve(x, h, 'full') = irfft(rfft(x, n) * rfft(h, n), n), n = N + L - 1,
    TA/functional/functional.py:2255-2258; crop keeps the shape, which is what
    /root/reference/tests/test_training_pipeline.py:242 asserts.  rir_idx < 0 = bypass."""
    x = np.asarray(x, dtype=dtype)
    out = x.copy()
    N = x.shape[1]
    for b, r in enumerate(np.asarray(rir_idx)):
        if r < 0:
            continue
        h = np.asarray(rirs[int(r)], dtype=np.float64)
        n = N + h.shape[0] - 1
        y = np.fft.irfft(np.fft.rfft(x[b].astype(np.float64), n) * np.fft.rfft(h, n), n)
        out[b] = y[:N].astype(dtype)
    return out


def gather_noise(bank: list[np.ndarray], noise_idx, noise_off, N: int, dtype=np.float64) -> np.ndarray:
    """Noise segment for each clip: bank[idx][(off + j) mod len], j < N (wraps, so a noise
    recording shorter than the clip loops).  Rows with idx < 0 are zero."""
    noise_idx = np.asarray(noise_idx)
    out = np.zeros((len(noise_idx), N), dtype=dtype)
    for b, i in enumerate(noise_idx):
        if i < 0:
            continue
        src = np.asarray(bank[int(i)])
        j = (int(noise_off[b]) + np.arange(N)) % src.shape[0]
        out[b] = src[j].astype(dtype)
    return out


def add_noise(x: np.ndarray, noise: np.ndarray, snr_db: np.ndarray, active=None) -> np.ndarray:
    """F.add_noise, TA/functional/functional.py:2374-2382:
    scale = 10 ** ((10 (log10 Es - log10 En) - snr) / 20), y = x + scale * noise."""
    dt = x.dtype.type
    es = (x.astype(x.dtype) ** 2).sum(axis=-1)
    en = (noise ** 2).sum(axis=-1)
    with np.errstate(divide="ignore", invalid="ignore"):
        snr0 = dt(10.0) * (np.log10(es) - np.log10(en))
        scale = dt(10.0) ** ((snr0 - np.asarray(snr_db, dtype=x.dtype)) / dt(20.0))
        y = x + scale[:, None] * noise
    if active is not None:
        y = np.where(np.asarray(active)[:, None], y, x)
    return y.astype(x.dtype)


# --------------------------------------------------------------------------------------
# A3: time-stretch / pitch-shift / resample
# --------------------------------------------------------------------------------------
PV_NFFT, PV_HOP = 512, 128       # F.pitch_shift defaults, TA/functional/functional.py:1596-1604


def stft_complex(x: np.ndarray, n_fft: int, hop_length: int, dtype=np.float64, window=None) -> np.ndarray:
    """torch.stft(center=True, reflect, periodic Hann, onesided, return_complex): (B, N) -> (B, K, T) complex."""
    x = np.asarray(x, dtype=dtype)
    pad = n_fft // 2
    xp = np.pad(x, ((0, 0), (pad, pad)), mode="reflect")
    T = num_frames(x.shape[1], hop_length)
    idx = (np.arange(T) * hop_length)[:, None] + np.arange(n_fft)[None, :]
    w = hann_periodic(n_fft, dtype) if window is None else np.asarray(window).astype(dtype)
    spec = np.fft.rfft((xp[:, idx] * w[None, None, :]).astype(np.float64), axis=-1)
    if dtype == np.float32:
        spec = spec.astype(np.complex64)
    return np.transpose(spec, (0, 2, 1))


def phase_vocoder(spec: np.ndarray, rate: float, hop_length: int, dtype=np.float64) -> np.ndarray:
    """F.phase_vocoder (TA/functional/functional.py:732-800): (B, K, T) -> (B, K, ceil(T / rate)).

    time_steps = arange(0, T, rate); frames idx = floor(time_steps) and idx + 1 (two zero frames appended);
    per bin the unwrapped phase increment angle1 - angle0 is accumulated (cumsum) starting from the phase of
    frame 0; magnitudes are interpolated linearly with alpha = time_steps mod 1."""
    if rate == 1.0:
        return spec
    B, K, T = spec.shape
    rdt = np.float32 if dtype == np.float32 else np.float64
    n_out = int(math.ceil(T / rate))
    ts = (np.arange(n_out, dtype=np.float64) * rate).astype(rdt)
    alphas = (ts % rdt(1.0)).astype(rdt)
    pa = _linspace32(0.0, math.pi * hop_length, K).astype(rdt)[None, :, None]   # float32 constants, see stretch_core
    sp = np.concatenate([spec, np.zeros((B, K, 2), spec.dtype)], axis=-1)
    i0 = ts.astype(np.int64)
    i1 = (ts + rdt(1.0)).astype(rdt).astype(np.int64)
    s0, s1 = sp[..., i0], sp[..., i1]
    a0, a1 = np.angle(s0).astype(rdt), np.angle(s1).astype(rdt)
    n0, n1 = np.abs(s0).astype(rdt), np.abs(s1).astype(rdt)
    two_pi = rdt(2.0 * math.pi)
    ph = (a1 - a0 - pa).astype(rdt)
    ph = (ph - two_pi * np.round(ph / two_pi)).astype(rdt)
    ph = (ph + pa).astype(rdt)
    ph = np.concatenate([np.angle(spec[..., :1]).astype(rdt), ph[..., :-1]], axis=-1)
    acc = np.cumsum(ph, axis=-1, dtype=rdt)
    mag = (alphas * n1 + (rdt(1.0) - alphas) * n0).astype(rdt)
    out = mag * np.cos(acc) + 1j * (mag * np.sin(acc))
    return out.astype(np.complex64 if dtype == np.float32 else np.complex128)


def istft(spec: np.ndarray, n_fft: int, hop_length: int, length: int, dtype=np.float64, window=None) -> np.ndarray:
    """torch.istft(center=True, window=periodic Hann, length=length): irfft per frame, * window, overlap-add,
    divide by the overlap-added squared window, drop n_fft/2 leading samples.  (B, K, T) -> (B, length)."""
    B, K, T = spec.shape
    w = (hann_periodic(n_fft, dtype) if window is None else np.asarray(window).astype(dtype)).astype(np.float64)
    frames = np.fft.irfft(np.transpose(spec, (0, 2, 1)).astype(np.complex128), n=n_fft, axis=-1)   # (B, T, n_fft)
    if dtype == np.float32:
        frames = frames.astype(np.float32).astype(np.float64)
    frames = frames * w
    total = n_fft + hop_length * (T - 1)
    y = np.zeros((B, total))
    env = np.zeros(total)
    for t in range(T):
        y[:, t * hop_length:t * hop_length + n_fft] += frames[:, t]
        env[t * hop_length:t * hop_length + n_fft] += w * w
    start = n_fft // 2
    if start + length > total:
        raise ValueError("istft: requested length exceeds the overlap-added signal")
    return (y[:, start:start + length] / env[start:start + length]).astype(dtype)


def stretch_core(x: np.ndarray, rate: float, dtype=np.float64, window=None) -> np.ndarray:
    """torchaudio's _stretch_waveform (TA/functional/functional.py:1644-1693) with the rate explicit:
    (B, N) -> (B, round(N / rate)).
    ``window`` may inject torch.hann_window(512)'s float32 values: they differ from the correctly rounded
    Hann window in the last 1-3 ulp, and a stretched PURE TONE is sensitive to exactly that (1e-4 relative,
    growing along the clip); noise-like signals are not."""
    N = x.shape[1]
    # _stretch_waveform builds window / phase_advance in float32 whatever the waveform dtype is
    w32 = hann_periodic(PV_NFFT, np.float32) if window is None else np.asarray(window, np.float32)
    spec = stft_complex(x, PV_NFFT, PV_HOP, dtype, window=w32)
    return istft(phase_vocoder(spec, rate, PV_HOP, dtype), PV_NFFT, PV_HOP, int(round(N / rate)), dtype, window=w32)


def _fix_len(y: np.ndarray, N: int) -> np.ndarray:
    """_fix_waveform_shape, TA/functional/functional.py:1696-1718."""
    return y[:, :N] if y.shape[1] >= N else np.pad(y, ((0, 0), (0, N - y.shape[1])))


def time_stretch(x: np.ndarray, rates, dtype=np.float64, window=None) -> np.ndarray:
    """Per clip stretch_core(rate) cropped / zero-padded back to N; rate == 1.0 = untouched.
    Reference surface: time_stretch_range, tests/test_training_pipeline.py:233."""
    out = np.asarray(x, dtype=dtype).copy()
    for b, r in enumerate([float(v) for v in rates]):
        if r != 1.0:
            out[b] = _fix_len(stretch_core(out[b:b + 1], r, dtype, window), x.shape[1])[0]
    return out


def sinc_resample_kernel(orig_freq: int, new_freq: int, dtype=np.float64, lowpass_filter_width: int = 6,
                         rolloff: float = 0.99):
    """_get_sinc_resample_kernel, 'sinc_interp_hann' (TA/functional/functional.py:1305-1402), frequencies
    already divided by their gcd - restricted to its support: torchaudio's dense kernel [new_freq][2*width +
    orig_freq] is zero (|t| clamped to the window's edge) except for the <= 2*width + 1 entries around
    q = phase * orig / new.  Returns (first [new_freq], taps [new_freq][2*width + 1], width) with
    taps[p][j] = kernel[p][first[p] + j] (0 where that index is outside the dense kernel).
    float32 follows torchaudio's float32 operation order (it builds the kernel in the waveform's dtype)."""
    f = np.float32 if dtype == np.float32 else np.float64
    base_freq = min(orig_freq, new_freq) * rolloff
    width = int(math.ceil(lowpass_filter_width * orig_freq / base_freq))
    ph = np.arange(new_freq)
    first = (ph * orig_freq) // new_freq                       # kernel column of the first tap (offset -width folded in)
    col = first[:, None] + np.arange(2 * width + 1)[None, :]   # dense-kernel column index = q + width
    q = col - width
    idx = (q.astype(f) / f(orig_freq)).astype(f)
    t = (((-ph).astype(f) / f(new_freq)).astype(f)[:, None] + idx).astype(f)
    t = (t * f(base_freq)).astype(f)
    t = np.clip(t, f(-lowpass_filter_width), f(lowpass_filter_width))
    window = np.cos(((t * f(math.pi)).astype(f) / f(lowpass_filter_width)).astype(f) / f(2)).astype(f) ** 2
    t = (t * f(math.pi)).astype(f)
    scale = f(base_freq / orig_freq)
    with np.errstate(invalid="ignore", divide="ignore"):
        k = np.where(t == 0, f(1.0), (np.sin(t).astype(f) / t).astype(f))
    k = (k * (window * scale).astype(f)).astype(f)
    k = np.where(col < 2 * width + orig_freq, k, f(0.0))
    return first, k, width


def resample(x: np.ndarray, orig_freq: int, new_freq: int, dtype=np.float64) -> np.ndarray:
    """F.resample (TA/functional/functional.py:1405-1497): zero-pad (width, width + orig), correlate with the
    polyphase kernel at stride orig, interleave the new_freq phases, keep ceil(new * N / orig) samples."""
    x = np.asarray(x, dtype=dtype)
    if orig_freq == new_freq:
        return x
    g = math.gcd(int(orig_freq), int(new_freq))
    o, n = int(orig_freq) // g, int(new_freq) // g
    first, taps, width = sinc_resample_kernel(o, n, dtype)
    B, N = x.shape
    nblk = N // o + 1
    xp = np.pad(x, ((0, 0), (width, width + o + 2 * width + 1))).astype(np.float64)
    target = int(np.ceil(np.float32(n * N / o)))             # torch.as_tensor(python float) is float32
    out = np.zeros((B, nblk * n))
    k64 = taps.astype(np.float64)
    cols = first[:, None] + np.arange(taps.shape[1])[None, :]
    for m in range(nblk):
        out[:, m * n:(m + 1) * n] = np.einsum("bpj,pj->bp", xp[:, m * o + cols], k64)
    return out[:, :target].astype(dtype)


def pitch_shift(x: np.ndarray, n_steps, sample_rate: int = 16000, dtype=np.float64, window=None) -> np.ndarray:
    """F.pitch_shift per clip (TA/functional/functional.py:1596-1641): stretch by rate = 2^(-n/12), resample
    int(sample_rate / rate) -> sample_rate, crop / zero-pad to N.  0 semitones = untouched."""
    out = np.asarray(x, dtype=dtype).copy()
    N = x.shape[1]
    for b, n in enumerate([int(v) for v in n_steps]):
        if n == 0:
            continue
        rate = 2.0 ** (-float(n) / 12)
        y = stretch_core(out[b:b + 1], rate, dtype, window)
        out[b] = _fix_len(resample(y, int(sample_rate / rate), sample_rate, dtype), N)[0]
    return out


def spec_mask(feat: np.ndarray, fstart=None, flen=None, tstart=None, tlen=None, mask_value=0.0) -> np.ndarray:
    """Explicit-index SpecAugment: rows [f0, f0+len) and columns [t0, t0+len) := mask_value.

    Index semantics of mask_along_axis(_iid), TA/functional/functional.py:864-870,939-953.
    feat: (B, ..., F, T); start/len arrays are (B, n_masks) integers."""
    out = feat.copy()
    B = feat.shape[0]
    F_, T_ = feat.shape[-2], feat.shape[-1]
    for b in range(B):
        if fstart is not None:
            for s, l in zip(np.atleast_1d(fstart[b]), np.atleast_1d(flen[b])):
                s = int(s); l = int(l)
                if l > 0:
                    out[b, ..., max(s, 0):min(s + l, F_), :] = mask_value
        if tstart is not None:
            for s, l in zip(np.atleast_1d(tstart[b]), np.atleast_1d(tlen[b])):
                s = int(s); l = int(l)
                if l > 0:
                    out[b, ..., :, max(s, 0):min(s + l, T_)] = mask_value
    return out


def features(x, *, sample_rate=16000, feature_type="mel", n_mels=128, n_mfcc=40, n_fft=1024,
             hop_length=160, top_db=80.0, use_cmvn=False, cmvn_eps=1e-5, dtype=np.float64, fb=None, window=None,
             dct=None):
    """FeatureExtractor.__call__ arithmetic on a batch: (B, N) -> (B, 1, F, T)."""
    if feature_type in ("mel", "mel_spectrogram"):
        f = log_mel(x, sample_rate, n_fft, hop_length, n_mels, top_db, dtype=dtype, fb=fb, window=window)
    elif feature_type == "mfcc":
        f = mfcc(x, sample_rate, n_fft, hop_length, n_mels, n_mfcc, top_db, dtype=dtype, dct=dct, fb=fb, window=window)
    else:
        raise ValueError(f"unknown feature_type {feature_type!r}")
    if use_cmvn:
        f = cmvn(f, cmvn_eps)
    return f[:, None, :, :]


def pipeline(x, *, rirs=None, rir_idx=None, noise_bank=None, noise_idx=None, noise_off=None,
             snr_db=None, fstart=None, flen=None, tstart=None, tlen=None, mask_value=0.0,
             dtype=np.float64, **feat_kw):
    """RIR reverb -> noise @ SNR -> features -> SpecAugment, all draws explicit.
    Order per SURVEY.md Appendix A ("RIR first, then noise")."""
    x = np.asarray(x, dtype=dtype)
    if rirs is not None and rir_idx is not None:
        x = rir_reverb(x, rirs, rir_idx, dtype)
    if noise_bank is not None and noise_idx is not None:
        nz = gather_noise(noise_bank, noise_idx, noise_off, x.shape[1], dtype)
        x = add_noise(x, nz, np.asarray(snr_db), active=np.asarray(noise_idx) >= 0)
    f = features(x, dtype=dtype, **feat_kw)
    if fstart is not None or tstart is not None:
        f = spec_mask(f, fstart, flen, tstart, tlen, mask_value)
    return f
