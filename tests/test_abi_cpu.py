"""CPU tests of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/msda_b200.h declares, argument validation works without touching a GPU, and the
Python module keeps the reference's surface (vision.cpp:4-7) and error behaviour."""
import ctypes
import os
import re

import pytest
import torch

from conftest import ROOT

HEADER = os.path.join(ROOT, "include", "msda_b200.h")


@pytest.fixture(scope="module")
def lib():
    from trackformer_b200 import _build
    return ctypes.CDLL(_build.build_library())


def declared_functions():
    names = set()
    inc = os.path.dirname(HEADER)
    for h in sorted(os.listdir(inc)):
        if h.endswith(".h"):
            src = re.sub(r"/\*.*?\*/", "", open(os.path.join(inc, h)).read(), flags=re.S)
            names |= set(re.findall(r"\b((?:msda_b200|tfb200)_\w+)\s*\(", src))
    return sorted(names)


def test_header_declares_expected_entry_points():
    names = declared_functions()
    for n in ("msda_b200_forward_f32", "msda_b200_forward_f64", "msda_b200_backward_f32",
              "msda_b200_backward_f64", "msda_b200_forward_host_f32", "msda_b200_backward_host_f32",
              "msda_b200_error_string", "msda_b200_abi_version"):
        assert n in names


def test_library_exports_every_declared_symbol(lib):
    for name in declared_functions():
        assert hasattr(lib, name), f"{name} declared in include/msda_b200.h but not exported"


def test_abi_version_matches_header(lib):
    v = int(re.search(r"#define\s+MSDA_B200_ABI_VERSION\s+(\d+)", open(HEADER).read()).group(1))
    assert lib.msda_b200_abi_version() == v


def test_argument_validation_without_gpu(lib):
    f = lib.msda_b200_forward_f32
    f.restype = ctypes.c_int
    vp, i = ctypes.c_void_p, ctypes.c_int
    f.argtypes = [vp] * 5 + [i] * 7 + [vp]
    lib.msda_b200_error_string.restype = ctypes.c_char_p
    lib.msda_b200_error_string.argtypes = [ctypes.c_int]
    buf = ctypes.create_string_buffer(64)
    p = ctypes.cast(buf, vp)
    assert f(p, p, p, p, p, 1, 0, 1, 4, 1, 1, 1, None) == -2          # S = 0 -> MSDA_E_DIMS
    assert f(p, p, p, p, p, 1, 4, 1, 4, 99, 1, 1, None) == -4         # L > MAX_LEVELS
    assert f(None, p, p, p, p, 1, 4, 1, 4, 1, 1, 1, None) == -1       # NULL value
    assert f(p, p, p, p, p, 1, 1 << 30, 8, 32, 1, 1, 1, None) == -3   # slab > 2^31-1 elements
    assert f(p, p, p, p, p, 0, 4, 1, 4, 1, 1, 1, None) == 0           # empty batch is a no-op
    assert b"NULL" in lib.msda_b200_error_string(-1)
    assert lib.msda_b200_error_string(0) == b"success"


def test_python_module_surface_and_cpu_rejection():
    from trackformer_b200 import _build, ext
    _build.build_all()
    m = ext.load()
    assert m.__name__ == "MultiScaleDeformableAttention"
    assert hasattr(m, "ms_deform_attn_forward") and hasattr(m, "ms_deform_attn_backward")
    value = torch.zeros(1, 4, 1, 4)
    shapes = torch.tensor([[2, 2]])
    loc = torch.zeros(1, 1, 1, 1, 1, 2)
    attn = torch.zeros(1, 1, 1, 1, 1)
    with pytest.raises(RuntimeError, match="Not implemented on the CPU"):   # ms_deform_attn.h:27
        m.ms_deform_attn_forward(value, shapes, loc, attn, 64)
    with pytest.raises(RuntimeError, match="Not implemented on the CPU"):   # ms_deform_attn.h:48
        m.ms_deform_attn_backward(value, shapes, loc, attn, torch.zeros(1, 1, 4), 64)


def test_function_has_no_cpu_fallback():
    from trackformer_b200.msda_function import MSDeformAttnFunction
    with pytest.raises(RuntimeError, match="Not implemented on the CPU"):
        MSDeformAttnFunction.apply(torch.zeros(1, 4, 1, 4), torch.tensor([[2, 2]]),
                                   torch.zeros(1, 1, 1, 1, 1, 2), torch.zeros(1, 1, 1, 1, 1), 64)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "trackformer_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
