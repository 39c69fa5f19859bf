"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/buffalo_b200.h declares, and refuses to run without a Blackwell GPU (no CPU fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "buffalo_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(bfl_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported():
    from buffalo_b200 import _cabi
    handle = ctypes.CDLL(_cabi.LIB_PATH)
    names = _declared()
    assert len(names) >= 40
    for n in names:
        assert hasattr(handle, n), "missing export %s" % n
    assert set(names) == set(_cabi.PROTOTYPES), set(names) ^ set(_cabi.PROTOTYPES)


def test_library_is_sm100_only():
    from buffalo_b200 import _cabi
    lib = _cabi.lib()
    assert lib.bfl_compiled_sm() == 100
    assert lib.bfl_abi_version() == 1


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from buffalo_b200 import _cabi, backend
    obj = backend.CuALS()
    with pytest.raises(_cabi.BackendError) as e:
        obj.init({"d": 16})
    assert "no CPU fallback" in str(e.value)
    sgd = backend.CuSGD("bpr")
    with pytest.raises(_cabi.BackendError):
        sgd.init({"d": 16})


def test_bad_option_file_returns_false_or_raises():
    # reference: init() returns False on a missing option file (lib/algo.cc:22-34) and the Python
    # layer asserts (buffalo/algo/als.py:43)
    from buffalo_b200 import backend
    obj = backend.CuALS()
    assert obj.init(b"/nonexistent/option.json") is False
    assert "File not exists" in obj.last_error


def test_product_never_imports_oracle():
    bad = []
    for pkg in ("buffalo_b200", "buffalo"):
        base = os.path.join(ROOT, pkg)
        for dp, _, fs in os.walk(base):
            for f in fs:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".sh")):
                    src = open(os.path.join(dp, f)).read()
                    if re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M) or "libbuffalo_oracle" in src:
                        bad.append(os.path.join(dp, f))
    assert not bad, bad
