"""The product library: built by nvcc for sm_100a, loads without a GPU, and
exports every symbol include/snfft.h declares.  No compute calls here."""
import ctypes
import os
import re
import subprocess

import pytest

from conftest import ROOT


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "snfft.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(snfft_[a-z_0-9]+)\s*\(", text)))


@pytest.fixture(scope="module")
def product_so():
    from algebraist_b200 import _build
    return _build.build()


def test_header_symbols_exported(product_so):
    lib = ctypes.CDLL(product_so)
    syms = _header_symbols()
    assert len(syms) >= 18
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/snfft.h but not exported"
    lib.snfft_build_info.restype = ctypes.c_char_p
    assert b"sm_100a" in lib.snfft_build_info()


def test_binding_matches_header(product_so):
    from algebraist_b200 import _lib
    assert sorted(_lib.SIGNATURES) == _header_symbols()
    _lib.bind(ctypes.CDLL(product_so))


def test_sass_is_sm100a(product_so):
    out = subprocess.run(["cuobjdump", "-lelf", product_so], capture_output=True, text=True)
    assert out.returncode == 0 and "sm_100a" in out.stdout


def test_no_cpu_fallback():
    import torch
    import algebraist_b200
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError):
        algebraist_b200.sn_fft(torch.randn(120), 5)


def test_product_does_not_import_oracle_or_emu():
    pkg = os.path.join(ROOT, "algebraist_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dirpath, fn)).read()
                assert "oracle" not in text.replace("no CPU or PyTorch fallback", ""), fn
                if fn.endswith(".py"):
                    assert "emu" not in text or fn == "plan.py", fn
