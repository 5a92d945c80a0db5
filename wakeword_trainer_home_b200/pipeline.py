"""Batched GPU feature pipeline: the thin PyTorch host layer over libwwfeat's C ABI.

``FeaturePlan`` owns one ``wwf_plan`` (device constants + registered noise / RIR banks) and
routes ``tensor.data_ptr()`` and the current CUDA stream into ``wwf_featurize`` /
``wwf_augment``.  PyTorch is used only for device memory and streams.

``AugParams`` is the explicit form of every random draw of the reference's
``AudioAugmentation`` / ``SpecAugment`` (BASELINE.json north_star: "RIR/noise indices, SNRs
and SpecAugment masks passed in explicitly"), so the same draws can be handed to the oracle.
"""
from __future__ import annotations

import os

import ctypes as C
from dataclasses import dataclass, fields
from typing import Optional, Sequence

import torch

from . import _native as N
from . import constants as K


def _dev_index(device) -> int:
    d = torch.device(device)
    if d.type != "cuda":
        raise N.WwfError(-3, f"device {device!r}: this package runs on CUDA (sm_100a) only; there is no CPU path")
    return d.index if d.index is not None else torch.cuda.current_device()


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


@dataclass
class AugParams:
    """Per-batch augmentation draws (all optional; ``None`` = that augmentation is off).

    rir_idx     int32 [B]      index into the registered RIR bank, < 0 = dry
    noise_idx   int32 [B]      index into the registered noise bank, < 0 = clean
    noise_off   int64 [B]      first sample inside the noise clip (wraps around its end)
    snr_db      float32 [B]    target SNR
    fmask_start/fmask_len  int32 [B, n_freq_masks]  masked feature rows [start, start+len)
    tmask_start/tmask_len  int32 [B, n_time_masks]  masked frames       [start, start+len)
    stretch_rate float64 [B]   time-stretch speed factor (> 1 = faster), exactly 1.0 = untouched
    pitch_steps int32 [B]      pitch shift in semitones, 0 = untouched
    stretch_lo / pitch_range   host-side bounds of the two arrays above (they size the workspace);
                               filled in automatically from host tensors or by the on-GPU draw
    all_reverb                 host-side fact about rir_idx, filled in the same way: True = every clip of the batch is
                               reverberated.  A scheduling hint only (FeaturePlan._conv_order): with dry clips in a
                               large batch the reverb kernel deals the reverberated ones round-robin over its CTAs,
                               which costs ~2 us that an all-reverb batch need not pay.  None = unknown (order on).
    Order of application: time-stretch -> pitch-shift -> reverb -> noise -> features -> masks.
    """
    rir_idx: Optional[torch.Tensor] = None
    noise_idx: Optional[torch.Tensor] = None
    noise_off: Optional[torch.Tensor] = None
    snr_db: Optional[torch.Tensor] = None
    fmask_start: Optional[torch.Tensor] = None
    fmask_len: Optional[torch.Tensor] = None
    tmask_start: Optional[torch.Tensor] = None
    tmask_len: Optional[torch.Tensor] = None
    stretch_rate: Optional[torch.Tensor] = None
    pitch_steps: Optional[torch.Tensor] = None
    stretch_lo: Optional[float] = None
    pitch_range: Optional[tuple] = None
    all_reverb: Optional[bool] = None
    resident_on: Optional[torch.device] = None   # set by .to(): every tensor already has the ABI's dtype on that device

    _DTYPES = dict(rir_idx=torch.int32, noise_idx=torch.int32, noise_off=torch.int64, snr_db=torch.float32,
                   fmask_start=torch.int32, fmask_len=torch.int32, tmask_start=torch.int32, tmask_len=torch.int32,
                   stretch_rate=torch.float64, pitch_steps=torch.int32)

    def tensor_fields(self):
        return [f.name for f in fields(self) if f.name in self._DTYPES]

    def to(self, device, non_blocking: bool = False) -> "AugParams":
        kw = dict(stretch_lo=self.stretch_lo, pitch_range=self.pitch_range, all_reverb=self.all_reverb)
        for name in self.tensor_fields():
            v = getattr(self, name)
            if v is not None:
                v = torch.as_tensor(v)
                # bounds of the shape-changing draws are taken while the values are still on the host
                if name == "stretch_rate" and kw["stretch_lo"] is None and not v.is_cuda and v.numel():
                    kw["stretch_lo"] = float(v.min())
                if name == "pitch_steps" and kw["pitch_range"] is None and not v.is_cuda and v.numel():
                    kw["pitch_range"] = (int(v.min()), int(v.max()))
                if name == "rir_idx" and kw["all_reverb"] is None and not v.is_cuda and v.numel():
                    kw["all_reverb"] = bool((v >= 0).all())
                v = v.to(device=device, dtype=self._DTYPES[name], non_blocking=non_blocking).contiguous()
            kw[name] = v
        moved = [kw[n] for n in self.tensor_fields() if kw.get(n) is not None]
        kw["resident_on"] = moved[0].device if moved else torch.device(device)    # "cuda" -> "cuda:0"
        return AugParams(**kw)

    def is_resident(self, device) -> bool:
        """Every tensor already lives on ``device`` with the ABI's dtype (nothing to convert before a call)."""
        if self.resident_on != device:
            return False
        for name in self.tensor_fields():
            v = getattr(self, name)
            if v is not None and (v.device != device or v.dtype != self._DTYPES[name] or not v.is_contiguous()):
                return False
        return True

    def nbytes(self) -> int:
        return sum(getattr(self, n).numel() * getattr(self, n).element_size()
                   for n in self.tensor_fields() if getattr(self, n) is not None)


@dataclass
class DrawConfig:
    """Distribution of the on-GPU augmentation draws (AugmentationConfig, src/config/defaults.py:73-95)."""
    seed: int = 0
    rir_prob: float = 0.25
    noise_prob: float = 0.5
    freq_mask_prob: float = 0.5
    time_mask_prob: float = 0.5
    snr_range: tuple = (5.0, 20.0)
    freq_mask_param: int = 15
    time_mask_param: int = 35
    stretch_prob: float = 0.0            # time_stretch_min/max, pitch_shift_min/max: src/config/defaults.py:76-79
    stretch_range: tuple = (0.8, 1.2)
    pitch_prob: float = 0.0
    pitch_range: tuple = (-2, 2)


def draw_mask_params(gen: Optional[torch.Generator], B: int, size: int, mask_param: int, n_masks: int, p: float = 1.0):
    """Integer (start, len) pairs from the two uniform draws torchaudio's mask_along_axis_iid makes
    (value = U*param, min_value = U*(size - value); start = floor(min_value), len = floor(value);
    torchaudio/functional/functional.py:806-810,864-870)."""
    if p != 1.0:
        mask_param = min(mask_param, int(size * p))
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


def _flatten_bank(clips: Sequence[torch.Tensor], device) -> tuple[torch.Tensor, list[int]]:
    offs = [0]
    for c in clips:
        if c.dim() != 1 or c.numel() == 0:
            raise ValueError("bank clips must be non-empty 1-D tensors")
        offs.append(offs[-1] + c.numel())
    flat = torch.cat([torch.as_tensor(c, dtype=torch.float32).reshape(-1) for c in clips]).to(device).contiguous()
    return flat, offs


class FeaturePlan:
    """One (device, feature config) plan.  Mirrors the reference's FeatureExtractor arguments
    (src/evaluation/evaluator.py:86-94) plus the knobs the fused path adds."""

    def __init__(self, sample_rate: int = 16000, feature_type: str = "mel", n_mels: int = 128, n_mfcc: int = 40,
                 n_fft: int = 1024, hop_length: int = 160, device="cuda", *, out_dtype=torch.float32,
                 cmvn: bool = False, top_db: Optional[float] = 80.0, f_min: float = 0.0, f_max: Optional[float] = None,
                 cmvn_eps: float = 1e-5, mask_value: float = 0.0, n_freq_masks: int = 0, n_time_masks: int = 0):
        if feature_type == "mel_spectrogram":            # legacy alias, evaluator.py:82-83
            feature_type = "mel"
        if feature_type not in ("mel", "mfcc"):
            raise ValueError(f"feature_type must be 'mel' or 'mfcc', got {feature_type!r}")
        if out_dtype not in (torch.float32, torch.float16):
            raise ValueError("out_dtype must be torch.float32 or torch.float16")
        self.lib = N.load()
        self.device = torch.device("cuda", _dev_index(device))
        self.sample_rate, self.feature_type = sample_rate, feature_type
        self.n_mels, self.n_mfcc, self.n_fft, self.hop_length = n_mels, n_mfcc, n_fft, hop_length
        self.out_dtype = out_dtype
        self.n_freq_masks, self.n_time_masks = n_freq_masks, n_time_masks
        f_max = float(sample_rate // 2) if f_max is None else float(f_max)

        # float32 constants bit-identical to torchaudio's (see constants.py)
        self._handle = C.c_void_p()
        window = mel = dct = None
        if n_fft in N.SUPPORTED_N_FFT and 0 < n_mels <= n_fft // 2 + 1:
            window = K.hann_window(n_fft)
            mel = K.mel_filterbank(n_fft // 2 + 1, f_min, f_max, n_mels, sample_rate)
            if feature_type == "mfcc" and 0 < n_mfcc <= n_mels:
                dct = K.dct_matrix(n_mfcc, n_mels)
        as_fp = lambda t: None if t is None else C.cast(t.data_ptr(), C.POINTER(C.c_float))
        cfg = N.Config(sample_rate=sample_rate, n_fft=n_fft, hop_length=hop_length, n_mels=n_mels, n_mfcc=n_mfcc,
                       feature_type=N.FEAT_MFCC if feature_type == "mfcc" else N.FEAT_LOGMEL,
                       out_dtype=N.OUT_F16 if out_dtype == torch.float16 else N.OUT_F32, cmvn=int(bool(cmvn)),
                       top_db=-1.0 if top_db is None else float(top_db), f_min=float(f_min), f_max=f_max,
                       cmvn_eps=float(cmvn_eps), mask_value=float(mask_value),
                       n_freq_masks=n_freq_masks, n_time_masks=n_time_masks,
                       window=as_fp(window), mel_fb=as_fp(mel), dct=as_fp(dct))
        N.check(self.lib.wwf_plan_create(C.byref(cfg), self.device.index, C.byref(self._handle)))
        info = N.Info()
        N.check(self.lib.wwf_plan_info(self._handle, C.byref(info)))
        self.n_feat, self.n_freq, self.sm_count = info.n_feat, info.n_freq, info.sm_count
        self._stretch_tables = False   # torch's float32 Hann window of the stretch stage is handed over on first use
        self._noise = None          # keeps the borrowed noise bank alive
        self._workspace: dict[int, torch.Tensor] = {}
        self._conv_order_pinned = bool(os.environ.get("WWF_NO_CONV_ORDER"))      # (the library read it at plan creation)
        self._conv_order_now = 0 if self._conv_order_pinned else 1
        self._ws_need: dict[tuple, int] = {}                    # wwf_workspace_bytes(B, n), a pure function of the plan
        self.n_noise = self.n_rir = 0

    # ------------------------------------------------------------------ lifetime
    def close(self):
        h, self._handle = getattr(self, "_handle", None), None
        if h:
            self.lib.wwf_plan_destroy(h)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ shape facts
    def num_frames(self, n_samples: int) -> int:
        """T = n_samples // hop + 1 (src/export/onnx_exporter.py:316-317)."""
        return n_samples // self.hop_length + 1

    def info(self) -> N.Info:
        info = N.Info()
        N.check(self.lib.wwf_plan_info(self._handle, C.byref(info)))
        return info

    def draw_aug(self, cfg: DrawConfig, first_index: int, B: int, n_samples: int) -> AugParams:
        """Augmentation draws for B clips made on the GPU (Philox4x32-10, counter = sample number
        first_index + i, key = cfg.seed): no host RNG, no H2D copy.  Returns device-resident AugParams
        ready for ``featurize``; tests/helpers.py:philox_draws recomputes them bit-exactly on the host."""
        dev = self.device
        a = AugParams(rir_idx=torch.empty(B, dtype=torch.int32, device=dev),
                      noise_idx=torch.empty(B, dtype=torch.int32, device=dev),
                      noise_off=torch.empty(B, dtype=torch.int64, device=dev),
                      snr_db=torch.empty(B, dtype=torch.float32, device=dev), resident_on=dev,
                      all_reverb=bool(cfg.rir_prob >= 1.0))
        if self.n_freq_masks:
            a.fmask_start = torch.empty(B, self.n_freq_masks, dtype=torch.int32, device=dev)
            a.fmask_len = torch.empty(B, self.n_freq_masks, dtype=torch.int32, device=dev)
        if self.n_time_masks:
            a.tmask_start = torch.empty(B, self.n_time_masks, dtype=torch.int32, device=dev)
            a.tmask_len = torch.empty(B, self.n_time_masks, dtype=torch.int32, device=dev)
        if cfg.stretch_prob > 0:
            a.stretch_rate = torch.empty(B, dtype=torch.float64, device=dev)
            a.stretch_lo = min(1.0, float(cfg.stretch_range[0]))
        if cfg.pitch_prob > 0:
            a.pitch_steps = torch.empty(B, dtype=torch.int32, device=dev)
            a.pitch_range = (min(0, int(cfg.pitch_range[0])), max(0, int(cfg.pitch_range[1])))
        dc = N.DrawConfig(seed=int(cfg.seed) & (2 ** 64 - 1), rir_prob=cfg.rir_prob, noise_prob=cfg.noise_prob,
                          freq_mask_prob=cfg.freq_mask_prob, time_mask_prob=cfg.time_mask_prob,
                          snr_lo=float(cfg.snr_range[0]), snr_hi=float(cfg.snr_range[1]),
                          freq_mask_param=int(cfg.freq_mask_param), time_mask_param=int(cfg.time_mask_param),
                          stretch_prob=cfg.stretch_prob, stretch_lo=float(cfg.stretch_range[0]), stretch_hi=float(cfg.stretch_range[1]),
                          pitch_prob=cfg.pitch_prob, pitch_lo=int(cfg.pitch_range[0]), pitch_hi=int(cfg.pitch_range[1]))
        st = N.Aug(*[_ptr(getattr(a, f[0])) for f in N.Aug._fields_])
        stream = torch.cuda.current_stream(dev)
        N.check(self.lib.wwf_draw_aug(self._handle, C.byref(dc), C.c_uint64(int(first_index)), B, self.num_frames(n_samples),
                                      C.byref(st), C.c_void_p(stream.cuda_stream)))
        return a

    def check_finite(self) -> bool:
        """True if every feature computed since the last call was finite (synchronises the current
        stream).  Mirrors the reference trainer's non-finite batch skip, src/training/trainer.py:177-179."""
        flag = C.c_int()
        stream = torch.cuda.current_stream(self.device)
        N.check(self.lib.wwf_check_finite(self._handle, C.c_void_p(stream.cuda_stream), C.byref(flag)))
        return flag.value == 0

    # ------------------------------------------------------------------ measurement hook
    def profile(self, enable: bool = True):
        N.check(self.lib.wwf_profile_enable(self._handle, int(enable)))

    def profile_read(self):
        """(avg reverb-kernel ms, avg feature-kernel ms, calls) since the last read."""
        c, f, n = C.c_double(), C.c_double(), C.c_int()
        N.check(self.lib.wwf_profile_read(self._handle, C.byref(c), C.byref(f), C.byref(n)))
        return c.value, f.value, n.value

    def profile_read_kernels(self):
        """({'conv_kernel', 'feat_prep_kernel', 'feat_frames_kernel' | 'feat_kernel', 'feat_epilogue_mma_kernel' |
        'feat_epilogue_block_kernel'} -> average ms per call, calls, calls that took the flat path) since the last read."""
        ms, n, ns = (C.c_double * 4)(), C.c_int(), C.c_int()
        N.check(self.lib.wwf_profile_read_kernels(self._handle, ms, C.byref(n), C.byref(ns)))
        split = n.value > 0 and ns.value == n.value
        names = ["conv_kernel", "feat_prep_kernel", "feat_frames_kernel" if split else "feat_kernel",
                 "feat_epilogue_mma_kernel" if self.feature_type == "mfcc" else "feat_epilogue_block_kernel"]
        return {k: v for k, v in zip(names, ms) if v > 0.0 or k == "conv_kernel"}, n.value, ns.value

    # ------------------------------------------------------------------ banks
    def _register(self, kind: int, clips: Sequence[torch.Tensor]):
        flat, offs = _flatten_bank(clips, self.device)
        arr = (C.c_int64 * len(offs))(*offs)
        stream = torch.cuda.current_stream(self.device)
        N.check(self.lib.wwf_bank_register(self._handle, kind, _ptr(flat), arr, len(clips), C.c_void_p(stream.cuda_stream)))
        self._ws_need.clear()                                   # the reverb workspace depends on the registered RIR bank
        return flat

    def register_noise(self, clips: Sequence[torch.Tensor]):
        """Background-noise recordings (data/raw/background in the reference, README.md:57)."""
        self._noise = self._register(N.BANK_NOISE, clips)   # borrowed by the plan -> keep alive
        self.n_noise = len(clips)

    def register_rirs(self, rirs: Sequence[torch.Tensor]):
        """Room impulse responses (data/raw/rirs, README.md:58); spectra are computed once here."""
        self._register(N.BANK_RIR, rirs)                     # consumed during the call
        self.n_rir = len(rirs)

    # ------------------------------------------------------------------ calls
    def _aug_struct(self, aug: Optional[AugParams], B: int):
        if aug is None:
            return None, None
        # draws made on the GPU (draw_aug) or moved once with .to(plan.device) are passed through untouched
        a = aug if aug.is_resident(self.device) else aug.to(self.device)
        for name in ("rir_idx", "noise_idx", "noise_off", "snr_db", "stretch_rate", "pitch_steps"):
            v = getattr(a, name)
            if v is not None and v.shape != (B,):
                raise ValueError(f"AugParams.{name} must have shape ({B},), got {tuple(v.shape)}")
        if a.noise_idx is not None and (a.noise_off is None or a.snr_db is None):
            raise ValueError("AugParams.noise_idx needs noise_off and snr_db")
        for s, l, n in (("fmask_start", "fmask_len", self.n_freq_masks), ("tmask_start", "tmask_len", self.n_time_masks)):
            vs, vl = getattr(a, s), getattr(a, l)
            if (vs is None) != (vl is None):
                raise ValueError(f"AugParams.{s} and {l} must be given together")
            if vs is not None and (vs.shape != (B, n) or vl.shape != (B, n)):
                raise ValueError(f"AugParams.{s}/{l} must have shape ({B}, {n}) for this plan, got {tuple(vs.shape)}")
        st = N.Aug(*[_ptr(getattr(a, f[0])) for f in N.Aug._fields_])
        return st, a            # keep `a` alive until the call returns

    def _ws(self, B: int, n: int, stream_id: int):
        need = self._ws_need.get((B, n))                        # (a ctypes round trip per call otherwise: the 1-clip path is host-bound)
        if need is None:
            if len(self._ws_need) >= 256:
                self._ws_need.clear()
            need = self._ws_need[(B, n)] = int(self.lib.wwf_workspace_bytes(self._handle, B, n))
        return self._ws_bytes(need, stream_id)

    def _ws_bytes(self, need: int, key):
        if need == 0:
            return None, 0
        ws = self._workspace.pop(key, None)
        if ws is None or ws.numel() < need:
            ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        self._workspace[key] = ws                               # most recently used last
        while len(self._workspace) > self.MAX_WORKSPACES:       # short-lived streams must not pin their buffers forever
            self._workspace.pop(next(iter(self._workspace)))
        return ws, need

    MAX_WORKSPACES = 8

    def release_workspaces(self):
        """Drop the cached per-stream scratch buffers (they are re-created on demand)."""
        self._workspace.clear()

    def set_path(self, path: str = "auto"):
        """Force the launch shape of ``featurize``: 'fused' (one kernel), 'flat' (flat frame queue + epilogue) or
        'auto' (the library picks per batch shape).  Test / measurement control (wwf_plan_set_option)."""
        code = {"auto": N.PATH_AUTO, "fused": N.PATH_FUSED, "flat": N.PATH_FLAT, "split": N.PATH_FLAT}[path]
        N.check(self.lib.wwf_plan_set_option(self._handle, N.OPT_FEAT_PATH, code))
        self._ws_need.clear()

    def set_pdl(self, enable: bool = True):
        """Chain the kernels of a call with programmatic dependent launch (default on)."""
        N.check(self.lib.wwf_plan_set_option(self._handle, N.OPT_PDL, int(bool(enable))))
        self._ws_need.clear()

    def _conv_order(self, aug: Optional[AugParams]):
        """Per batch: WWF_OPT_CONV_ORDER off iff the draws are known to reverberate every clip (AugParams.all_reverb)."""
        want = 0 if (aug is not None and aug.all_reverb) else 1
        if want != self._conv_order_now and not self._conv_order_pinned:
            N.check(self.lib.wwf_plan_set_option(self._handle, N.OPT_CONV_ORDER, want))
            self._conv_order_now = want

    def set_conv_order(self, enable: Optional[bool] = None):
        """Reverb work items dealt round-robin over the CTAs (True), in batch order (False), or chosen per batch from
        ``AugParams.all_reverb`` (None, the default).  Test / measurement control."""
        self._conv_order_pinned = enable is not None
        if enable is not None:
            N.check(self.lib.wwf_plan_set_option(self._handle, N.OPT_CONV_ORDER, int(bool(enable))))
            self._conv_order_now = int(bool(enable))

    def set_epilogue_warp(self, enable: bool = True):
        """MFCC calls without SpecAugment flags of the common shapes take the warp-autonomous tensor-core epilogue
        (default on); off = the block-wise one for every call.  Test / measurement control."""
        N.check(self.lib.wwf_plan_set_option(self._handle, N.OPT_EPILOGUE_WARP, int(bool(enable))))
        self._ws_need.clear()

    # ------------------------------------------------------------------ waveform-shape augmentations
    def time_stretch(self, wav: torch.Tensor, rates: torch.Tensor, rate_lo: Optional[float] = None,
                     out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """(B, N) -> (B, N): clip b played rates[b] times faster without changing its pitch (torchaudio's
        STFT -> phase_vocoder -> iSTFT), cropped / zero-padded to N; rates[b] == 1.0 copies the clip.
        ``rate_lo`` is a lower bound of the rates (sizes the workspace); it is read from ``rates`` when omitted."""
        with self._range("wwf.time_stretch"):
            return self._time_stretch(wav, rates, rate_lo, out)

    def _time_stretch(self, wav, rates, rate_lo, out):
        wav = self._check_wav(wav)
        B, n = wav.shape
        rates = torch.as_tensor(rates)
        if rates.shape != (B,):
            raise ValueError(f"rates must have shape ({B},), got {tuple(rates.shape)}")
        if rate_lo is None:
            rate_lo = float(rates.min())                       # synchronises if the rates live on the GPU
        rates = rates.to(device=self.device, dtype=torch.float64).contiguous()
        if out is None:
            out = torch.empty(B, n, dtype=torch.float32, device=self.device)
        stream = torch.cuda.current_stream(self.device)
        rate_lo = min(1.0, float(rate_lo))
        self._ensure_stretch_tables()
        ws, ws_bytes = self._ws_bytes(int(self.lib.wwf_stretch_workspace_bytes(B, n, rate_lo)), (stream.cuda_stream, "pv"))
        N.check(self.lib.wwf_time_stretch(self._handle, _ptr(wav), B, n, wav.stride(0), _ptr(rates), rate_lo,
                                          _ptr(out), out.stride(0), _ptr(ws), ws_bytes, C.c_void_p(stream.cuda_stream)))
        return out

    def pitch_shift(self, wav: torch.Tensor, n_steps: torch.Tensor, step_range: Optional[tuple] = None,
                    out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """(B, N) -> (B, N): torchaudio F.pitch_shift by n_steps[b] semitones per clip (0 copies the clip).
        ``step_range`` = (lo, hi) bounds of n_steps; read from the tensor when omitted."""
        with self._range("wwf.pitch_shift"):
            return self._pitch_shift(wav, n_steps, step_range, out)

    def _pitch_shift(self, wav, n_steps, step_range, out):
        wav = self._check_wav(wav)
        B, n = wav.shape
        n_steps = torch.as_tensor(n_steps)
        if n_steps.shape != (B,):
            raise ValueError(f"n_steps must have shape ({B},), got {tuple(n_steps.shape)}")
        if step_range is None:
            step_range = (int(n_steps.min()), int(n_steps.max()))
        lo, hi = min(0, int(step_range[0])), max(0, int(step_range[1]))
        n_steps = n_steps.to(device=self.device, dtype=torch.int32).contiguous()
        if out is None:
            out = torch.empty(B, n, dtype=torch.float32, device=self.device)
        stream = torch.cuda.current_stream(self.device)
        self._ensure_stretch_tables()
        ws, ws_bytes = self._ws_bytes(int(self.lib.wwf_pitch_workspace_bytes(B, n, lo, hi)), (stream.cuda_stream, "pv"))
        N.check(self.lib.wwf_pitch_shift(self._handle, _ptr(wav), B, n, wav.stride(0), _ptr(n_steps), lo, hi,
                                         _ptr(out), out.stride(0), _ptr(ws), ws_bytes, C.c_void_p(stream.cuda_stream)))
        return out

    def resample(self, wav: torch.Tensor, orig_freq: int, new_freq: int, n_out: Optional[int] = None) -> torch.Tensor:
        """(B, n) at orig_freq -> (B, ceil(new * n / orig)) at new_freq, torchaudio F.resample defaults
        (the AudioProcessor step in front of the path, src/evaluation/evaluator.py:76-79)."""
        wav = self._check_wav(wav)
        B, n = wav.shape
        if n_out is None:
            n_out = int(self.lib.wwf_resample_length(n, int(orig_freq), int(new_freq)))
        out = torch.empty(B, n_out, dtype=torch.float32, device=self.device)
        stream = torch.cuda.current_stream(self.device)
        N.check(self.lib.wwf_resample(self._handle, _ptr(wav), B, n, wav.stride(0), int(orig_freq), int(new_freq),
                                      _ptr(out), n_out, out.stride(0), C.c_void_p(stream.cuda_stream)))
        return out

    def _ensure_stretch_tables(self):
        if not self._stretch_tables:
            w = K.hann_window(512)
            N.check(self.lib.wwf_set_stretch_window(self._handle, C.cast(w.data_ptr(), C.POINTER(C.c_float))))
            self._stretch_tables = True

    def _shape_augs(self, wav: torch.Tensor, a: Optional[AugParams]) -> torch.Tensor:
        """time-stretch, then pitch-shift, when the draws ask for them (out of place)."""
        if a is None:
            return wav
        if a.stretch_rate is not None:
            wav = self.time_stretch(wav, a.stretch_rate, a.stretch_lo)
        if a.pitch_steps is not None:
            wav = self.pitch_shift(wav, a.pitch_steps, a.pitch_range)
        return wav

    def _check_wav(self, wav: torch.Tensor) -> torch.Tensor:
        if wav.dim() != 2:
            raise ValueError(f"expected a (B, N) batch, got shape {tuple(wav.shape)}")
        if wav.device != self.device or wav.dtype != torch.float32:
            wav = wav.to(device=self.device, dtype=torch.float32)
        if wav.stride(1) != 1:
            wav = wav.contiguous()
        return wav

    def nvtx(self, enable: bool = True):
        """Bracket every ``featurize`` / ``augment`` / ``time_stretch`` / ``pitch_shift`` call with an NVTX range
        (``wwf.featurize`` ...) so the path shows up by stage on an Nsight timeline (profiling)."""
        self._nvtx = bool(enable)

    def _range(self, name: str):
        if getattr(self, "_nvtx", False):
            return torch.cuda.nvtx.range(name)
        import contextlib
        return contextlib.nullcontext()

    def featurize(self, wav: torch.Tensor, aug: Optional[AugParams] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """(B, N) float32 clips -> (B, 1, n_feat, T) features, one fused GPU pass
        (reverb -> noise -> STFT -> mel -> dB -> [DCT] -> [CMVN] -> [masks])."""
        if not getattr(self, "_nvtx", False):
            return self._featurize(wav, aug, out)
        with self._range("wwf.featurize"):
            return self._featurize(wav, aug, out)

    def _featurize(self, wav, aug, out):
        wav = self._check_wav(wav)
        B, n = wav.shape
        T = self.num_frames(n)
        if out is None:
            out = torch.empty(B, 1, self.n_feat, T, dtype=self.out_dtype, device=self.device)
        elif out.shape != (B, 1, self.n_feat, T) or out.dtype != self.out_dtype or out.device != self.device or not out.is_contiguous():
            raise ValueError("out must be a contiguous (B, 1, n_feat, T) tensor of the plan's dtype on the plan's device")
        st, keep = self._aug_struct(aug, B)
        self._conv_order(keep)
        wav = self._shape_augs(wav, keep)
        stream = torch.cuda.current_stream(self.device)
        ws, ws_bytes = self._ws(B, n, stream.cuda_stream)
        N.check(self.lib.wwf_featurize(self._handle, _ptr(wav), B, n, wav.stride(0), None if st is None else C.byref(st),
                                       _ptr(out), self.n_feat * T, _ptr(ws), ws_bytes, C.c_void_p(stream.cuda_stream)))
        return out

    def augment(self, wav: torch.Tensor, aug: Optional[AugParams], out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Time-domain half only: (B, N) -> (B, N) float32 ([time-stretch] -> [pitch-shift] -> reverb -> noise @ SNR)."""
        wav = self._check_wav(wav)
        B, n = wav.shape
        if out is None:
            out = torch.empty(B, n, dtype=torch.float32, device=self.device)
        st, keep = self._aug_struct(aug, B)
        self._conv_order(keep)
        wav = self._shape_augs(wav, keep)
        stream = torch.cuda.current_stream(self.device)
        ws, ws_bytes = self._ws(B, n, stream.cuda_stream)
        N.check(self.lib.wwf_augment(self._handle, _ptr(wav), B, n, wav.stride(0), None if st is None else C.byref(st),
                                     _ptr(out), out.stride(0), _ptr(ws), ws_bytes, C.c_void_p(stream.cuda_stream)))
        return out


def as_sequence(features: torch.Tensor) -> torch.Tensor:
    """(B, 1, F, T) features -> the (B, T, F) view the reference's LSTM / GRU classifiers take
    (src/models/architectures.py:177 "Input tensor (batch, time_steps, features)").  A view: no copy; call
    ``.contiguous()`` where a kernel needs it (cuDNN RNNs accept the strided view)."""
    if features.dim() != 4 or features.shape[1] != 1:
        raise ValueError(f"expected (B, 1, F, T) features, got {tuple(features.shape)}")
    return features.squeeze(1).transpose(1, 2)


def gather_clips(bank: torch.Tensor, idx: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out[b] = bank[idx[b]] as float32 from a device-resident clip bank (n, N): float32, or int16 PCM
    (converted as x / 32768).  The batch-assembly step of the device-resident loader."""
    if not bank.is_cuda or bank.dim() != 2 or bank.stride(1) != 1 or bank.dtype not in (torch.float32, torch.int16):
        raise ValueError("bank must be a CUDA (n, N) float32 or int16 tensor with contiguous rows")
    idx = idx.to(device=bank.device, dtype=torch.int64).contiguous()
    B, n = idx.numel(), bank.shape[1]
    if out is None:
        out = torch.empty(B, n, dtype=torch.float32, device=bank.device)
    stream = torch.cuda.current_stream(bank.device)
    N.check(N.load().wwf_gather_clips(_ptr(bank), N.BANK_I16 if bank.dtype == torch.int16 else N.BANK_F32, bank.shape[0], n,
                                      bank.stride(0), _ptr(idx), B, _ptr(out), out.stride(0), bank.device.index,
                                      C.c_void_p(stream.cuda_stream)))
    return out


def peak_normalize(wav: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Per-clip ``x / max|x|`` of a CUDA float32 (N,) or (B, N) tensor (all-zero clips unchanged) -
    the normalisation the reference applies to a chunk before FeatureExtractor
    (src/evaluation/inference.py:189-191).  ``out=wav`` normalises in place."""
    if not wav.is_cuda or wav.dtype != torch.float32 or wav.dim() not in (1, 2):
        raise ValueError("wav must be a CUDA float32 (N,) or (B, N) tensor")
    w2 = wav.reshape(-1, wav.shape[-1])
    if w2.stride(1) != 1:
        w2 = w2.contiguous()
    o2 = torch.empty_like(w2) if out is None else out.reshape(-1, wav.shape[-1])
    stream = torch.cuda.current_stream(wav.device)
    N.check(N.load().wwf_peak_normalize(_ptr(w2), w2.shape[0], w2.shape[1], w2.stride(0), _ptr(o2), o2.stride(0),
                                        wav.device.index, C.c_void_p(stream.cuda_stream)))
    return o2.reshape(wav.shape)


def spec_augment_(spec: torch.Tensor, fmask_start=None, fmask_len=None, tmask_start=None, tmask_len=None,
                  mask_value: float = 0.0) -> torch.Tensor:
    """In-place explicit-index SpecAugment of a contiguous CUDA (B, F, T) float32/float16 tensor."""
    if spec.dim() != 3 or not spec.is_cuda or not spec.is_contiguous() or spec.dtype not in (torch.float32, torch.float16):
        raise ValueError("spec must be a contiguous CUDA (B, F, T) float32/float16 tensor")
    lib = N.load()
    B, F_, T_ = spec.shape
    dev = spec.device

    def prep(s, l):
        if s is None:
            return None, None, 0
        s = torch.as_tensor(s).to(dev, torch.int32).reshape(B, -1).contiguous()
        l = torch.as_tensor(l).to(dev, torch.int32).reshape(B, -1).contiguous()
        return s, l, s.shape[1]

    fs, fl, nf = prep(fmask_start, fmask_len)
    ts, tl, nt = prep(tmask_start, tmask_len)
    stream = torch.cuda.current_stream(dev)
    N.check(lib.wwf_spec_augment(_ptr(spec), N.OUT_F16 if spec.dtype == torch.float16 else N.OUT_F32, B, F_, T_, F_ * T_,
                                 _ptr(fs), _ptr(fl), nf, _ptr(ts), _ptr(tl), nt, float(mask_value),
                                 dev.index, C.c_void_p(stream.cuda_stream)))
    return spec
