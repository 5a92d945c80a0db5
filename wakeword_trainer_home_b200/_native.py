"""ctypes binding of libwwfeat.so (C ABI in include/wwfeat.h).

The library is built in-tree by ``__graft_entry__.build()`` /
``python -m wakeword_trainer_home_b200.build`` into ``wakeword_trainer_home_b200/lib/``.
There is no fallback: if the shared object is missing or a call fails, a
``WwfError`` is raised.
"""
from __future__ import annotations

import ctypes as C
import os

LIB_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib")
LIB_PATH = os.environ.get("WWF_LIB", os.path.join(LIB_DIR, "libwwfeat.so"))   # WWF_LIB: A/B-test another build

WWF_OK = 0
FEAT_LOGMEL, FEAT_MFCC = 0, 1
OUT_F32, OUT_F16 = 0, 1
BANK_NOISE, BANK_RIR = 0, 1
BANK_F32, BANK_I16 = 0, 1
OPT_FEAT_PATH, OPT_PDL, OPT_EPILOGUE_WARP, OPT_CONV_ORDER = 0, 1, 2, 3
PATH_AUTO, PATH_FUSED, PATH_FLAT = 0, 1, 2
MAX_MASKS = 8
SUPPORTED_N_FFT = (256, 400, 512, 1024, 2048)


class WwfError(RuntimeError):
    """A libwwfeat call failed (status code + wwf_last_error message)."""

    def __init__(self, code: int, msg: str):
        super().__init__(f"libwwfeat error {code}: {msg}")
        self.code = code


_fp = C.POINTER(C.c_float)


class Config(C.Structure):
    _fields_ = [
        ("sample_rate", C.c_int32), ("n_fft", C.c_int32), ("hop_length", C.c_int32),
        ("n_mels", C.c_int32), ("n_mfcc", C.c_int32), ("feature_type", C.c_int32),
        ("out_dtype", C.c_int32), ("cmvn", C.c_int32),
        ("top_db", C.c_float), ("f_min", C.c_float), ("f_max", C.c_float),
        ("cmvn_eps", C.c_float), ("mask_value", C.c_float),
        ("n_freq_masks", C.c_int32), ("n_time_masks", C.c_int32),
        ("window", _fp), ("mel_fb", _fp), ("dct", _fp),
    ]


class Aug(C.Structure):
    _fields_ = [
        ("rir_idx", C.c_void_p), ("noise_idx", C.c_void_p), ("noise_off", C.c_void_p),
        ("snr_db", C.c_void_p), ("fmask_start", C.c_void_p), ("fmask_len", C.c_void_p),
        ("tmask_start", C.c_void_p), ("tmask_len", C.c_void_p),
        ("stretch_rate", C.c_void_p), ("pitch_steps", C.c_void_p),
    ]


class DrawConfig(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("rir_prob", C.c_double), ("noise_prob", C.c_double),
                ("freq_mask_prob", C.c_double), ("time_mask_prob", C.c_double),
                ("snr_lo", C.c_float), ("snr_hi", C.c_float),
                ("freq_mask_param", C.c_int32), ("time_mask_param", C.c_int32),
                ("stretch_prob", C.c_double), ("stretch_lo", C.c_double), ("stretch_hi", C.c_double),
                ("pitch_prob", C.c_double), ("pitch_lo", C.c_int32), ("pitch_hi", C.c_int32)]


class Info(C.Structure):
    _fields_ = [
        ("n_freq", C.c_int32), ("n_feat", C.c_int32), ("device", C.c_int32), ("sm_count", C.c_int32),
        ("rir_fft_size", C.c_int32), ("rir_max_len", C.c_int32), ("n_rir", C.c_int32), ("n_noise", C.c_int32),
    ]


# every symbol include/wwfeat.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "wwf_version": (C.c_int, []),
    "wwf_last_error": (C.c_char_p, []),
    "wwf_plan_create": (C.c_int, [C.POINTER(Config), C.c_int, C.POINTER(C.c_void_p)]),
    "wwf_plan_destroy": (None, [C.c_void_p]),
    "wwf_plan_info": (C.c_int, [C.c_void_p, C.POINTER(Info)]),
    "wwf_num_frames": (C.c_int, [C.c_void_p, C.c_int]),
    "wwf_bank_register": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.POINTER(C.c_int64), C.c_int, C.c_void_p]),
    "wwf_workspace_bytes": (C.c_size_t, [C.c_void_p, C.c_int, C.c_int]),
    "wwf_featurize": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int64, C.POINTER(Aug),
                                C.c_void_p, C.c_int64, C.c_void_p, C.c_size_t, C.c_void_p]),
    "wwf_augment": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int64, C.POINTER(Aug),
                              C.c_void_p, C.c_int64, C.c_void_p, C.c_size_t, C.c_void_p]),
    "wwf_gather_clips": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_int, C.c_int64, C.c_void_p, C.c_int, C.c_void_p,
                                   C.c_int64, C.c_int, C.c_void_p]),
    "wwf_draw_aug": (C.c_int, [C.c_void_p, C.POINTER(DrawConfig), C.c_uint64, C.c_int, C.c_int, C.POINTER(Aug), C.c_void_p]),
    "wwf_stretch_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_double]),
    "wwf_set_stretch_window": (C.c_int, [C.c_void_p, _fp]),
    "wwf_time_stretch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int64, C.c_void_p, C.c_double,
                                   C.c_void_p, C.c_int64, C.c_void_p, C.c_size_t, C.c_void_p]),
    "wwf_pitch_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int, C.c_int]),
    "wwf_pitch_shift": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int64, C.c_void_p, C.c_int, C.c_int,
                                  C.c_void_p, C.c_int64, C.c_void_p, C.c_size_t, C.c_void_p]),
    "wwf_resample_length": (C.c_int, [C.c_int, C.c_int, C.c_int]),
    "wwf_resample": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int64, C.c_int, C.c_int,
                               C.c_void_p, C.c_int, C.c_int64, C.c_void_p]),
    "wwf_peak_normalize": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int64, C.c_void_p, C.c_int64, C.c_int, C.c_void_p]),
    "wwf_spec_augment": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int64,
                                   C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                                   C.c_float, C.c_int, C.c_void_p]),
    "wwf_check_finite": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_int)]),
    "wwf_profile_enable": (C.c_int, [C.c_void_p, C.c_int]),
    "wwf_profile_read": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int)]),
    "wwf_profile_read_kernels": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "wwf_plan_set_option": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "wwf_launch_count": (C.c_int64, []),
    "wwf_debug_poison_smem": (C.c_int, [C.c_int, C.c_uint32]),
}

_lib = None


def load() -> C.CDLL:
    """Load libwwfeat.so (once) and type every exported symbol.  Raises if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise WwfError(-3, f"{LIB_PATH} is not built - run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(nvcc, sm_100a).  There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)      # AttributeError if the .so does not export it
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != WWF_OK:
        raise WwfError(rc, load().wwf_last_error().decode("utf-8", "replace"))


def launch_count() -> int:
    return int(load().wwf_launch_count())
