import contextlib
import ctypes
import math
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
EMU_DIR = os.path.join(ROOT, "tests", "emu")
EMU_SO = os.path.join(EMU_DIR, "_build", "libsnfft_emu.so")
CSRC = os.path.join(ROOT, "algebraist_b200", "csrc")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def key(lam):
    return "_".join(str(x) for x in lam)


@pytest.fixture(scope="session")
def golden_tables():
    return np.load(os.path.join(GOLDEN, "tables.npz"))


def load_transform_golden(n):
    return np.load(os.path.join(GOLDEN, f"transform_n{n}.npz"))


def build_emu():
    """g++ build of the CUDA sources against tests/emu/cuda_emu.h (tests only).  Same translation units as the
    product build (algebraist_b200/_build.py), objects cached under tests/emu/_build/."""
    from concurrent.futures import ThreadPoolExecutor
    from algebraist_b200 import _build as PB
    bdir = os.path.dirname(EMU_SO)
    os.makedirs(bdir, exist_ok=True)
    PB.generate_col_headers()
    emu_h = os.path.join(EMU_DIR, "cuda_emu.h")
    units = [(os.path.join(bdir, os.path.basename(obj)), src, extra, deps + [emu_h]) for obj, src, extra, deps in PB.units()]
    todo = [u for u in units if PB.stale(u[0], u[3])]

    def compile_one(u):
        obj, src, extra, _ = u
        subprocess.run(["g++", "-std=c++17", "-O1", "-x", "c++", "-DSNFFT_HOST_EMU", "-I" + EMU_DIR, "-fPIC"] + extra +
                       ["-c", "-o", obj, src], check=True)

    if todo:
        with ThreadPoolExecutor(max_workers=max(1, min(len(todo), os.cpu_count() or 1))) as ex:
            list(ex.map(compile_one, todo))
    if todo or not os.path.exists(EMU_SO):
        subprocess.run(["g++", "-shared", "-o", EMU_SO] + [u[0] for u in units], check=True)
    return EMU_SO


@pytest.fixture(scope="session")
def emu():
    """The C ABI compiled for the host (kernel logic check without a GPU)."""
    from algebraist_b200 import _lib
    return _lib.bind(ctypes.CDLL(build_emu()))


class CAbi:
    """Thin numpy front end over the C ABI, for a library whose 'device' pointers
    are host pointers (the emulation build)."""

    def __init__(self, lib):
        self.lib = lib

    def plan(self, n, dtype):
        h = ctypes.c_void_p()
        rc = self.lib.snfft_plan_create(n, 0 if dtype == np.float32 else 1, 0, ctypes.byref(h))
        assert rc == 0, self.lib.snfft_last_error()
        return h

    def forward(self, plan, f):
        f = np.ascontiguousarray(f)
        B = f.shape[0]
        out = np.zeros(f.size, dtype=f.dtype)
        ws = np.zeros(max(self.lib.snfft_workspace_bytes(plan, B), 8), dtype=np.uint8)
        rc = self.lib.snfft_forward(plan, f.ctypes.data, B, out.ctypes.data, ws.ctypes.data, ws.nbytes, None)
        assert rc == 0, self.lib.snfft_last_error()
        return out

    def inverse(self, plan, packed, B):
        packed = np.ascontiguousarray(packed)
        out = np.zeros((B, packed.size // B), dtype=packed.dtype)
        ws = np.zeros(max(self.lib.snfft_workspace_bytes(plan, B), 8), dtype=np.uint8)
        rc = self.lib.snfft_inverse(plan, packed.ctypes.data, B, out.ctypes.data, ws.ctypes.data, ws.nbytes, None)
        assert rc == 0, self.lib.snfft_last_error()
        return out


def unpack(packed, n, B):
    """packed C-ABI coefficient array -> dict lambda -> (B, d, d)."""
    from oracle import sn_oracle as O
    out, off = {}, 0
    for lam in O.generate_partitions(n):
        d = O.dim(lam)
        out[lam] = packed[off * B:(off + d * d) * B].reshape(B, d, d)
        off += d * d
    assert off == math.factorial(n)
    return out


def pack(ft, n, B, dtype):
    from oracle import sn_oracle as O
    return np.concatenate([np.asarray(ft[lam], dtype=dtype).reshape(-1) for lam in O.generate_partitions(n)])


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


@contextlib.contextmanager
def package_on_emu(emu_lib):
    """Route the Python package through the emulation library with CPU tensors.
    Test-only: checks the wrappers' shape/dtype contract without a GPU."""
    import torch
    from algebraist_b200 import _lib, fourier, plan
    saved = (_lib._lib, plan._DEVICE_TYPES, fourier._to_device, fourier._stream_ptr, fourier._device_ctx,
             dict(plan._cache))
    _lib._lib = emu_lib
    plan._DEVICE_TYPES = ("cuda", "cpu")
    plan._cache.clear()
    fourier._to_device = lambda t: (t, t.device)
    fourier._stream_ptr = lambda device: None
    fourier._device_ctx = lambda device: contextlib.nullcontext()
    try:
        yield
    finally:
        plan._cache.clear()
        (_lib._lib, plan._DEVICE_TYPES, fourier._to_device, fourier._stream_ptr, fourier._device_ctx, cache) = saved
        plan._cache.update(cache)
