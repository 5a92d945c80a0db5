"""Mount this package as the reference's missing ``src.data`` package.

The reference imports ``from src.data.feature_extraction import FeatureExtractor``, ``from src.data.audio_utils import
AudioProcessor`` (src/evaluation/evaluator.py:14-15, src/evaluation/inference.py:20-21), ``from src.data.augmentation
import AudioAugmentation, SpecAugment`` (tests/test_training_pipeline.py:21) and ``from src.data.dataset import
WakewordDataset, load_dataset_splits`` (src/ui/panel_evaluation.py:28, src/ui/panel_training.py) - but never committed
the package.  ``install_as_src_data()`` registers modules of those names that re-export the
B200 implementations, so the reference's evaluator / trainer / tests import and run unchanged:

    import wakeword_trainer_home_b200.compat as compat
    compat.install_as_src_data()            # before the first `import src.evaluation...`
    from src.evaluation.evaluator import ModelEvaluator

A maintainer would instead commit four two-line files under ``src/data/``; this function is the same
thing done at run time, used by the tests and by ``bench.py --config cfg3``.
"""
from __future__ import annotations

import importlib
import sys
import types

_MODULES = {
    "feature_extraction": ("wakeword_trainer_home_b200.feature_extraction", ("FeatureExtractor",)),
    "augmentation": ("wakeword_trainer_home_b200.augmentation", ("AudioAugmentation", "SpecAugment")),
    "audio_utils": ("wakeword_trainer_home_b200.audio_utils", ("AudioProcessor",)),
    "dataset": ("wakeword_trainer_home_b200.dataset", ("WakewordDataset", "load_dataset_splits")),
}


def install_as_src_data(force: bool = False) -> None:
    """Register ``src.data`` and its four submodules in ``sys.modules`` (no-op if a real ``src.data`` is importable,
    unless ``force``).  ``src`` itself must be importable (the reference checkout or its pip-installed copy) or is
    created as an empty namespace package."""
    if not force:
        try:
            importlib.import_module("src.data.feature_extraction")
            return
        except Exception:
            pass
    try:
        src = importlib.import_module("src")
    except Exception:
        src = types.ModuleType("src")
        src.__path__ = []          # namespace-like: lets `import src.x` look at sys.modules first
        sys.modules["src"] = src
    pkg = types.ModuleType("src.data")
    pkg.__path__ = []
    pkg.__doc__ = "B200-native replacement of the reference's src.data package (wakeword_trainer_home_b200)"
    sys.modules["src.data"] = pkg
    setattr(src, "data", pkg)
    for name, (target, symbols) in _MODULES.items():
        real = importlib.import_module(target)
        mod = types.ModuleType(f"src.data.{name}")
        for s in symbols:
            setattr(mod, s, getattr(real, s))
            setattr(pkg, s, getattr(real, s))
        mod.__doc__ = real.__doc__
        sys.modules[f"src.data.{name}"] = mod
        setattr(pkg, name, mod)


def reference_on_path() -> bool:
    """Make the reference's ``src`` package importable: /root/reference in the build container, else the copy that
    ``pip install --target baseline/_ref`` left next to the repo (it travels to the GPU box).  True if found."""
    import os
    try:
        importlib.import_module("src.config.defaults")
        return True
    except Exception:
        pass
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for cand in (os.path.join(root, "baseline", "_ref"), "/root/reference"):
        if os.path.isdir(os.path.join(cand, "src", "config")):
            sys.path.insert(0, cand)
            sys.modules.pop("src", None)
            try:
                importlib.import_module("src.config.defaults")
                return True
            except Exception:
                sys.path.remove(cand)
    return False
