"""The C-ABI shared library: loads without a GPU, exports every symbol that
include/mgb.h declares, and its descriptor validation / workspace queries (pure host
code) behave.  No compute calls here."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def declared_symbols():
    text = (ROOT / "include" / "mgb.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mgb_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_binding_table():
    from mygrad_b200 import _lib

    assert declared_symbols() == sorted(_lib.EXPORTS)


def test_library_exports_every_declared_symbol():
    from mygrad_b200 import _lib

    handle = C.CDLL(str(_lib.LIB_PATH))
    for name in declared_symbols():
        assert hasattr(handle, name), f"{name} declared in include/mgb.h but not exported"
    assert b"sm_100a" in _lib.lib.mgb_version()


def test_library_contains_sm100a_code_only():
    import shutil
    import subprocess

    from mygrad_b200 import _lib

    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not Path(cuobjdump).exists():
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "-lelf", str(_lib.LIB_PATH)], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def _conv_desc(dtype, N, C_, F, X, W, s, p, d):
    from mygrad_b200._lib import ConvDesc

    desc = ConvDesc()
    desc.dtype, desc.nd, desc.N, desc.C, desc.F = dtype, len(X), N, C_, F
    for i in range(len(X)):
        desc.X[i], desc.W[i], desc.stride[i], desc.pad[i], desc.dil[i] = X[i], W[i], s[i], p[i], d[i]
    return desc


def test_workspace_query_and_validation_run_on_host():
    from mygrad_b200._lib import lib

    n = C.c_size_t(0)
    d = _conv_desc(1, 256, 64, 128, (56, 56), (3, 3), (1, 1), (1, 1), (1, 1))
    assert lib.mgb_conv_workspace_bytes(C.byref(d), 0, C.byref(n)) == 0
    # f64: one 16-byte tap entry per (c, k), then the packed filters [F][tap][C] of the TMA-fed DMMA kernel
    assert n.value == 64 * 9 * 16 + 128 * 9 * 64 * 8 + 1024       # + alignment slack for the TMA source
    d = _conv_desc(0, 256, 64, 128, (56, 56), (3, 3), (1, 1), (1, 1), (1, 1))
    assert lib.mgb_conv_workspace_bytes(C.byref(d), 0, C.byref(n)) == 0
    assert n.value >= 2 * 256 * 56 * 56 * 64 * 4       # f32/tcgen05: channels-last hi+lo planes of x
    assert lib.mgb_conv_workspace_bytes(C.byref(d), 1, C.byref(n)) == 0 and n.value > 0
    assert lib.mgb_conv_workspace_bytes(C.byref(d), 2, C.byref(n)) == 0 and n.value > 0
    # BASELINE config 3 as written is invalid in the reference too (conv.py:63-74)
    bad = _conv_desc(1, 128, 128, 256, (64, 64), (5, 5), (2, 2), (0, 0), (2, 2))
    assert lib.mgb_conv_workspace_bytes(C.byref(bad), 0, C.byref(n)) != 0
    assert b"incompatible" in lib.mgb_last_error()
    ok = _conv_desc(1, 128, 128, 256, (65, 65), (5, 5), (2, 2), (0, 0), (2, 2))
    assert lib.mgb_conv_workspace_bytes(C.byref(ok), 1, C.byref(n)) == 0
    assert lib.mgb_conv_workspace_bytes(C.byref(d), 7, C.byref(n)) != 0
    d.dtype = 5
    assert lib.mgb_conv_workspace_bytes(C.byref(d), 0, C.byref(n)) != 0


def test_no_cpu_fallback_without_device():
    import mygrad_b200 as mg

    if mg.device_available():
        pytest.skip("a GPU is present")
    with pytest.raises(RuntimeError):
        mg.Tensor(np.ones((2, 2)))
    with pytest.raises(RuntimeError):
        mg.empty((4,), np.float32)


def test_product_never_imports_the_oracle():
    pkg = ROOT / "mygrad_b200"
    for f in pkg.rglob("*.py"):
        text = f.read_text()
        assert "oracle" not in re.sub(r'""".*?"""', "", text, flags=re.S).replace("# ", ""), f
    for f in (pkg / "csrc").glob("*.cu*"):
        assert "oracle" not in f.read_text(), f
