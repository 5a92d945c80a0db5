"""CPU oracle for the audio feature / augmentation hot path.

THIS PACKAGE IS TEST INFRASTRUCTURE.  It is the checker, never the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it.  The product package
(``wakeword_trainer_home_b200``) never imports anything from here and has no
CPU fallback.

What it restates
----------------
The reference's hot path lives in ``src/data/{feature_extraction,augmentation}.py``
which is *absent* from the reference checkout (git-ignored), so
the arithmetic is defined by the third-party dependency those files call:
**torchaudio, pinned 2.1.2+cu118** at ``/root/reference/requirements.txt:6``
(this image ships torchaudio 2.11.0; the formulas used here are unchanged).

Two layers:

* ``ta_oracle``  - the reconstructed ``FeatureExtractor`` / ``AudioAugmentation`` /
  ``SpecAugment`` call surface (call sites
  ``src/evaluation/evaluator.py:86-94,125``, ``src/evaluation/inference.py:94-102,197``,
  ``tests/test_training_pipeline.py:230-262``) implemented by calling torchaudio's
  own CPU ops with every random draw passed in explicitly.
* ``np_oracle``  - an independent numpy (float64 or float32) restatement of the same
  published formulas, used to pin ``ta_oracle`` and to keep the
  checker alive on a box without torchaudio.

Parity pinning
--------------
The reference's own tests hold NO golden vectors / KATs for this path (only shape
and finiteness asserts, ``tests/test_training_pipeline.py:242,243,262``), and the
path's module cannot be imported, so parity is **unpinned by the reference's own
tests**.  It is pinned instead to torchaudio itself: ``tests/golden/*.npz`` were
produced by ``tests/golden/make_golden.py`` calling torchaudio's transforms
directly in the build container, and both oracle layers are checked against them.
"""
from . import np_oracle  # noqa: F401

try:  # torchaudio is in the image; keep the numpy layer usable without it
    from . import ta_oracle  # noqa: F401
    HAVE_TORCHAUDIO = True
except Exception:  # pragma: no cover
    ta_oracle = None
    HAVE_TORCHAUDIO = False
