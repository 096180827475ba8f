"""Oracle (TEST INFRASTRUCTURE): import the real reference in place, in the authoring container.

``/root/reference`` exists only in the authoring container, never on the GPU box.  This loader
is used by ``tests/golden/make_golden.py`` to produce the committed golden vectors, and by CPU
tests that are skipped when the reference tree is absent.  It never copies reference sources.

Two modes:
  * :func:`load_hot_modules` pre-registers empty ``yolox`` package stubs so that
    ``yolox.models.embedding`` / ``activation`` / ``yolox.data.datasets.gen1`` import without
    running ``yolox/models/__init__.py`` (which needs spikingjelly).
  * :func:`load_full_model` puts ``oracle/sj_shim`` (spikingjelly + thop stand-ins) and the
    reference on ``sys.path`` so ``EventExp.get_model()`` runs unmodified.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

REF = os.environ.get("EAS_REFERENCE", "/root/reference")
_HERE = os.path.dirname(os.path.abspath(__file__))


def available() -> bool:
    return os.path.isdir(os.path.join(REF, "yolox"))


def load_hot_modules():
    """Returns (embedding, activation, gen1) reference modules."""
    R = os.path.join(REF, "yolox")
    if "yolox" not in sys.modules or not hasattr(sys.modules["yolox"], "__eas_stub__"):
        for name, path in [("yolox", R), ("yolox.models", R + "/models"), ("yolox.utils", R + "/utils"),
                           ("yolox.data", R + "/data"), ("yolox.data.datasets", R + "/data/datasets"),
                           ("yolox.utils.psee_loader", R + "/utils/psee_loader"),
                           ("yolox.utils.psee_loader.io", R + "/utils/psee_loader/io")]:
            m = types.ModuleType(name)
            m.__path__ = [path]
            m.__eas_stub__ = True
            sys.modules[name] = m
    emb = importlib.import_module("yolox.models.embedding")
    act = importlib.import_module("yolox.models.activation")
    gen1 = importlib.import_module("yolox.data.datasets.gen1")
    return emb, act, gen1


def make_ref_dataset(gen1, H, W, Tm):
    ds = object.__new__(gen1.GEN1Dataset)
    ds.img_size = (H, W)
    ds.slice_args = {"micro_slice": Tm}
    return ds


def load_full_model(name="e-yolox-s", opts=()):
    """Build the reference model through its own ``get_exp``; needs a clean ``yolox`` import."""
    if hasattr(sys.modules.get("yolox"), "__eas_stub__"):   # drop the stub packages of load_hot_modules
        for k in [k for k in sys.modules if k == "yolox" or k.startswith("yolox.")]:
            del sys.modules[k]
    root = os.path.dirname(_HERE)
    # the real spikingjelly wins when it is installed; the shim (our restatement) is only the stand-in
    try:
        import spikingjelly.activation_based.neuron  # noqa: F401
        paths = (root, REF)
    except ImportError:
        paths = (root, os.path.join(_HERE, "sj_shim"), REF)
    for p in paths:
        if p not in sys.path:
            sys.path.insert(0, p)
    from yolox.exp import get_exp  # type: ignore
    exp = get_exp(None, name)
    exp.merge(list(opts))
    model = exp.get_model()
    return exp, model


def spikingjelly_origin() -> str:
    """'shim' when the neuron / BN / container classes come from oracle/sj_shim, else the real package's version."""
    import spikingjelly
    f = getattr(spikingjelly, "__file__", "") or ""
    if "sj_shim" in f:
        return "shim"
    return "spikingjelly " + str(getattr(spikingjelly, "__version__", "unknown"))
