"""CPU: the C-ABI boundary — libovk.so loads, exports every symbol include/ovk.h declares, and the ctypes prototype
table covers the header both ways.  No compute calls (there is no GPU here)."""
import ctypes
import os
import re

import pytest

from openvision_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "ovk.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ovk_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_entry_points():
    syms = declared_symbols()
    for must in ("ovk_gemm_bf16", "ovk_layernorm_fwd", "ovk_attention_fwd", "ovk_version", "ovk_last_error"):
        assert must in syms


def test_library_exports_every_declared_symbol():
    assert os.path.exists(_lib.LIB_PATH), "libovk.so not built: python -c 'import __graft_entry__ as g; g.build()'"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, f"declared in include/ovk.h but not exported by libovk.so: {missing}"


def test_ctypes_table_matches_header():
    assert sorted(_lib.exported_symbols()) == declared_symbols()


def test_version_and_error_string_without_gpu():
    lib = _lib.load()
    assert lib.ovk_version() == 100 or lib.ovk_version() > 100
    # no CUDA device here: the support query must fail with an error code and a message, not crash
    rc = lib.ovk_device_supported()
    import torch
    if not torch.cuda.is_available():
        assert rc != 0
        assert lib.ovk_last_error()


def test_no_header_mentions_torch_types():
    src = open(HEADER).read()
    assert "at::" not in src and "torch::" not in src and "#include <torch" not in src


def test_missing_library_fails_loudly(tmp_path, monkeypatch):
    monkeypatch.setattr(_lib, "_lib", None)
    with pytest.raises(_lib.OvkError):
        _lib.load(str(tmp_path / "nope.so"))
    monkeypatch.setattr(_lib, "_lib", None)
