"""wakeword_trainer_home_b200 - B200 (sm_100a) audio feature / augmentation path.

Drop-in for the ``src.data`` feature-extraction and augmentation surface of
sarpel/wakeword_trainer_home, executed by hand-written CUDA kernels behind a C ABI
(include/wwfeat.h -> lib/libwwfeat.so).  No CPU fallback: constructing any class needs the
built library and a CUDA device.
"""
from ._native import WwfError, LIB_PATH, launch_count  # noqa: F401
from .pipeline import (AugParams, DrawConfig, FeaturePlan, as_sequence, draw_mask_params, gather_clips,  # noqa: F401
                       peak_normalize, spec_augment_)
from .feature_extraction import FeatureExtractor  # noqa: F401
from .augmentation import AudioAugmentation, SpecAugment  # noqa: F401
from .loader import DeviceBatchLoader, GpuBatchLoader, StreamedFeaturizer  # noqa: F401
from .sharding import shard_range, shard_seed  # noqa: F401
from .audio_utils import AudioProcessor, read_wav  # noqa: F401
from .formats import NpyFeatureLoader, load_npy, load_split_manifest, precompute_features, save_split_manifest  # noqa: F401
from .dataset import WakewordDataset, load_dataset_splits  # noqa: F401
from . import config_adapter  # noqa: F401
from .config_adapter import draw_config_from, plan_from_config  # noqa: F401

__version__ = "0.2.0"
__all__ = ["FeatureExtractor", "AudioAugmentation", "SpecAugment", "FeaturePlan", "AugParams",
           "draw_mask_params", "as_sequence", "spec_augment_", "peak_normalize", "WwfError", "launch_count", "GpuBatchLoader", "DeviceBatchLoader", "StreamedFeaturizer", "DrawConfig", "gather_clips",
           "shard_range", "shard_seed", "AudioProcessor", "read_wav", "NpyFeatureLoader", "load_npy", "load_split_manifest",
           "precompute_features", "save_split_manifest", "WakewordDataset", "load_dataset_splits", "config_adapter",
           "plan_from_config", "draw_config_from"]
