"""The C-ABI library loads and exports every symbol include/biogo_align.h declares; without a
GPU it fails loudly instead of falling back to anything."""
import ctypes as C
import os
import re

import pytest

import helpers
from biogo_b200 import _lib


def _declared_symbols():
    hdr = open(os.path.join(helpers.ROOT, "include", "biogo_align.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(ba_[a-z_]+)\s*\(", hdr)))


def test_header_and_binding_agree():
    assert _declared_symbols() == sorted(_lib.ABI_SYMBOLS)


def test_cuda_library_exports_every_declared_symbol(built):
    lib = C.CDLL(_lib.DEFAULT_LIB)
    for sym in _declared_symbols():
        assert hasattr(lib, sym), sym
    l = _lib.load()
    assert l.ba_abi_version() == 3
    assert b"sm_100a" in l.ba_build_info()


def test_no_cpu_fallback_without_a_gpu(built):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(_lib.AlignError) as e:
        _lib.Context(0)
    assert e.value.code == _lib.BA_ERR_CUDA


def test_product_loader_never_points_at_the_emulator():
    assert "emu" not in _lib.DEFAULT_LIB
    assert os.path.basename(_lib.DEFAULT_LIB) == "libbiogoalign.so"
    src = open(os.path.join(helpers.ROOT, "biogo_b200", "_lib.py")).read() + \
        open(os.path.join(helpers.ROOT, "biogo_b200", "align.py")).read()
    assert "oracle" not in src.replace("no CPU fallback", "")
