"""CPU-side checks of the drop-in boundary: the C-ABI library loads without a GPU and exports
every symbol include/xq_b200.h declares; compute entry points are NOT called here."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "xq_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(xq_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def libpath():
    import xq_native
    if not os.path.exists(xq_native.LIB_PATH):
        xq_native.build()
    return xq_native.LIB_PATH


def test_library_exports_every_declared_symbol(libpath):
    L = ctypes.CDLL(libpath)
    syms = header_symbols()
    assert len(syms) >= 10
    for s in syms:
        assert hasattr(L, s), f"{s} declared in include/xq_b200.h but not exported"


def test_binding_covers_header(libpath):
    import xq_native
    xq_native.lib()
    assert set(header_symbols()) <= set(xq_native.EXPORTS)


def test_no_cpu_fallback(libpath):
    """Without a GPU the engine must fail loudly, not compute on the host."""
    import torch
    import xq_native
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(xq_native.XqError):
        xq_native.Engine(0)
    # and the C entry point itself reports the missing device
    h = ctypes.c_void_p()
    rc = xq_native.lib().xq_create(0, ctypes.byref(h))
    assert rc < 0
    assert b"no CUDA device" in xq_native.lib().xq_last_error(None)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "xiangqi-alphazero_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f), errors="ignore").read()
                assert "xq_oracle" not in src and "oracle/" not in src, f
