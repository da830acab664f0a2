"""CPU checks of the drop-in boundary: the C-ABI library loads and exports every symbol of include/gibbs_b200.h."""
import ctypes
import os
import re

from conftest import ROOT


def _declared():
    src = open(os.path.join(ROOT, "include", "gibbs_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gibbs_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported():
    import __graft_entry__ as ge
    ge.build()
    lib = ctypes.CDLL(os.path.join(ROOT, "lda_thesis_b200", "libgibbs_b200.so"))
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), "include/gibbs_b200.h declares %s but the library does not export it" % n


def test_binding_declares_every_symbol(gibbs):
    lib = gibbs.load_library()
    for n in _declared():
        fn = getattr(lib, n)
        assert fn.argtypes is not None or n in ("gibbs_last_error", "gibbs_version", "gibbs_device_count"), n


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "lda_thesis_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(import|from)\s+(oracle|oracle_lib|patched_reference|philox)\b", text, re.M), f
                assert "liboracle" not in text, f
