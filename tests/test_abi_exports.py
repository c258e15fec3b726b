"""The C-ABI library loads on a machine without a GPU and exports every symbol include/cnhe.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "cnhe.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cnhe_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported_and_bound():
    from cryptonets_b200 import _lib
    L = _lib.lib()
    names = _declared()
    assert len(names) > 50
    for n in names:
        assert hasattr(L, n), n
    assert sorted(_lib.EXPORTS) == names  # the Python binding covers the whole header


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from cryptonets_b200.engine import Engine
    from cryptonets_b200 import CnheError
    with pytest.raises(CnheError, match="no CPU fallback"):
        Engine([40961], 4096)


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "cryptonets_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle_py" not in text and "bfv_oracle" not in text and "liboracle" not in text, f


def test_csharp_shim_matches_the_header_and_the_reference_interfaces():
    """integration/B200Native.cs (the shim a maintainer adds next to HE Wrapper/IFactory.cs) cannot be compiled here; its [DllImport]
    block must at least agree with include/cnhe.h in names, arity and marshalled types, and the three classes must carry every
    member of IVector / IMatrix / IFactory / IComputationEnvironment."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("check_csharp_bindings", os.path.join(ROOT, "tools", "check_csharp_bindings.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    errors, n = mod.check()
    assert not errors, errors
    assert n >= 78
