"""Import shims that let the UNMODIFIED reference modules run in this container -- TEST INFRASTRUCTURE.

The reference (``/root/reference``) imports four packages that are not installed here and cannot
be fetched: ``vector_quantize_pytorch`` and ``pytorch_model_summary`` (models/vq_brain.py:6-7),
``simple_parsing`` (models/brainformer.py:11, models/simple_mae:11) and ``accelerate``
(utils/train_utils.py:7,9).  The stubs below stand in for them:

* ``vector_quantize_pytorch.VectorQuantize`` -> ``oracle.vector_quantize_ref.VectorQuantizeRef``
  (the fp32 restatement; parity unpinned, see that file's header)
* ``pytorch_model_summary.summary``          -> no-op
* ``simple_parsing.helpers.Serializable``    -> empty base class

``load_reference()`` returns the reference's own ``models.vq_brain``, ``models.brainformer`` and
``models/simple_mae`` (the latter has no ``.py`` suffix, hence the SourceFileLoader).

This only works where ``/root/reference`` exists (the build container).  It is used by
``scripts/make_golden.py`` to produce the fixtures under ``tests/golden`` -- nothing that runs on
the GPU box imports it.
"""
from __future__ import annotations

import importlib
import importlib.machinery
import importlib.util
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("FK_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "models"))


def install_stubs() -> None:
    from oracle.vector_quantize_ref import VectorQuantizeRef

    if "vector_quantize_pytorch" not in sys.modules:
        m = types.ModuleType("vector_quantize_pytorch")
        m.VectorQuantize = VectorQuantizeRef
        m.ResidualVQ = None  # imported at models/vq_brain.py:6 but never used
        sys.modules["vector_quantize_pytorch"] = m
    if "pytorch_model_summary" not in sys.modules:
        m = types.ModuleType("pytorch_model_summary")
        m.summary = lambda *a, **k: None
        sys.modules["pytorch_model_summary"] = m
    if "simple_parsing" not in sys.modules:
        sp = types.ModuleType("simple_parsing")
        helpers = types.ModuleType("simple_parsing.helpers")

        class Serializable:  # the reference only uses it as a dataclass base
            pass

        helpers.Serializable = Serializable
        sp.helpers = helpers
        sp.ArgumentParser = object
        sys.modules["simple_parsing"] = sp
        sys.modules["simple_parsing.helpers"] = helpers


def load_reference():
    """Returns (vq_brain, brainformer, simple_mae) modules of the unmodified reference."""
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    install_stubs()
    out = []
    for name, fname in (("_ref_vq_brain", "models/vq_brain.py"),
                        ("_ref_brainformer", "models/brainformer.py"),
                        ("_ref_simple_mae", "models/simple_mae")):
        if name in sys.modules:
            out.append(sys.modules[name])
            continue
        path = os.path.join(REFERENCE_ROOT, fname)
        loader = importlib.machinery.SourceFileLoader(name, path)
        spec = importlib.util.spec_from_loader(name, loader)
        mod = importlib.util.module_from_spec(spec)
        sys.modules[name] = mod
        loader.exec_module(mod)
        out.append(mod)
    return tuple(out)
