"""TEST INFRASTRUCTURE ONLY (never imported by the product path).

Loads the *unmodified* reference module
``/root/reference/saicinpainting/training/modules/ffc.py`` under a private name so
that golden vectors can be generated from it (tests/golden/make_golden.py) and the
restatements in ``oracle/`` can be pinned against it.

The reference lives only in the build container: ``/root/reference`` does not exist
on the GPU box, so nothing that runs there (``-m gpu`` tests, ``smoke()``,
``bench.py``) may call :func:`load_reference_ffc`; ``available()`` says whether the
tree is present.

Two third-party imports of the reference are absent from this image and are not on
the numeric path: ``kornia.geometry.transform.rotate`` (used only by
``LearnableSpatialTransformWrapper``, spatial_transform.py:4) and
``pytorch_lightning.seed_everything`` (saicinpainting/utils.py:12).  They are
replaced by inert stubs.
"""
import importlib
import importlib.util
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("LAMA_REFERENCE_ROOT", "/root/reference")
_FFC_REL = "saicinpainting/training/modules/ffc.py"
_cached = None


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, _FFC_REL))


def _stub(name, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)
    sys.modules[name] = mod
    return mod


def _install_stubs():
    def _unavailable(*_a, **_k):
        raise RuntimeError("stubbed third-party symbol called on the oracle path")

    try:
        import kornia  # noqa: F401
    except Exception:
        k = _stub("kornia")
        g = _stub("kornia.geometry")
        t = _stub("kornia.geometry.transform", rotate=_unavailable)
        k.geometry = g
        g.transform = t
    try:
        import pytorch_lightning  # noqa: F401
    except Exception:
        _stub("pytorch_lightning", seed_everything=_unavailable)


def load_reference_ffc():
    """Return the reference ``ffc`` module object (classes FourierUnit ... FFCResNetGenerator)."""
    global _cached
    if _cached is not None:
        return _cached
    if not available():
        raise FileNotFoundError(f"reference tree not present at {REFERENCE_ROOT}")
    _install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.append(REFERENCE_ROOT)
    # ffc.py:10-13 imports its siblings through the package; let those resolve normally,
    # then load ffc.py itself under a private name so a drop-in registered as
    # ``saicinpainting.training.modules.ffc`` (lama_b200.patch) is never confused with it.
    spec = importlib.util.spec_from_file_location(
        "_lama_reference_ffc", os.path.join(REFERENCE_ROOT, _FFC_REL))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    _cached = mod
    return mod
