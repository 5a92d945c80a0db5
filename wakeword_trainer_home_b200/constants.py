"""Host-side float32 constants of the feature path, computed with the same torch ops (and in
the same order) torchaudio uses, so they are bit-identical to the oracle's.

Why here and not only in C: the triangular mel weights divide by differences of ``f_pts``
that are one powf away from each other - a 1-ulp difference in ``powf`` moves a weight by
~1e-5 relative (measured against libm).  Computing them with torch's own kernels removes
that source of drift; libwwfeat still has its own (libm) builders for callers without torch.
"""
from __future__ import annotations

import math
import torch


def hann_window(n_fft: int) -> torch.Tensor:
    """torch.hann_window(n_fft) (periodic) - Spectrogram's default window,
    torchaudio/transforms/_transforms.py:64-77."""
    return torch.hann_window(n_fft, periodic=True, dtype=torch.float32)


def mel_filterbank(n_freqs: int, f_min: float, f_max: float, n_mels: int, sample_rate: int) -> torch.Tensor:
    """HTK triangular filterbank, norm=None, (n_freqs, n_mels) float32.
    Same op sequence as torchaudio/functional/functional.py:490-587 (melscale_fbanks)."""
    all_freqs = torch.linspace(0, sample_rate // 2, n_freqs)
    m_min = 2595.0 * math.log10(1.0 + (f_min / 700.0))
    m_max = 2595.0 * math.log10(1.0 + (f_max / 700.0))
    m_pts = torch.linspace(m_min, m_max, n_mels + 2)
    f_pts = 700.0 * (10.0 ** (m_pts / 2595.0) - 1.0)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)
    down_slopes = (-1.0 * slopes[:, :-2]) / f_diff[:-1]
    up_slopes = slopes[:, 2:] / f_diff[1:]
    return torch.max(torch.zeros(1), torch.min(down_slopes, up_slopes)).contiguous()


def dct_matrix(n_mfcc: int, n_mels: int) -> torch.Tensor:
    """Orthonormal DCT-II, (n_mels, n_mfcc) float32 - torchaudio/functional/functional.py:636-667."""
    n = torch.arange(float(n_mels))
    k = torch.arange(float(n_mfcc)).unsqueeze(1)
    dct = torch.cos(math.pi / float(n_mels) * (n + 0.5) * k)
    dct[0] *= 1.0 / math.sqrt(2.0)
    dct *= math.sqrt(2.0 / float(n_mels))
    return dct.t().contiguous()

