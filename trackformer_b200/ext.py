"""Loader of the in-tree native extension.

``load()`` returns the compiled module ``MultiScaleDeformableAttention`` (same name and entry
points as the reference's extension, src/trackformer/models/ops/src/vision.cpp:4-7).  There is
no fallback of any kind: if the shared objects are missing the import raises with the build
command, and the op itself rejects CPU tensors ("Not implemented on the CPU", like the
reference's src/ms_deform_attn.h:27).
"""
from __future__ import annotations

import importlib.util
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
_NAME = "MultiScaleDeformableAttention"
_mod = None


def extension_path() -> str:
    return os.path.join(_HERE, _NAME + ".so")


def library_path() -> str:
    return os.path.join(_HERE, "libmsda_b200.so")


def load():
    """Import (once) and return the compiled extension module."""
    global _mod
    if _mod is not None:
        return _mod
    path = extension_path()
    if not (os.path.exists(path) and os.path.exists(library_path())):
        raise ImportError(
            f"trackformer_b200: native extension not built ({path} missing). "
            f"Run `python -c 'import __graft_entry__ as g; g.build()'` or `python trackformer_b200/_build.py`. "
            f"There is no CPU/PyTorch fallback for the MSDeformAttn hot path.")
    import torch  # noqa: F401  (libtorch must be loaded before the extension)
    if _NAME in sys.modules and getattr(sys.modules[_NAME], "__file__", None) == path:
        _mod = sys.modules[_NAME]
        return _mod
    spec = importlib.util.spec_from_file_location(_NAME, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    # registered under the reference's module name so `import MultiScaleDeformableAttention` works too
    sys.modules.setdefault(_NAME, mod)
    _mod = mod
    return mod
