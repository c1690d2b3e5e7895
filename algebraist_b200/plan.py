"""Per-(n, dtype, device) plan cache: the tables the reference rebuilds on
every call (SnIrrep construction, irreps.py:40-155; generate_partitions,
tableau.py:103-123) are built once by `snfft_plan_create` and kept."""
import ctypes
import math
import threading

import torch

from . import _lib

_DTYPES = {torch.float32: 0, torch.float64: 1}
_cache = {}
_cache_lock = threading.Lock()

MAX_N = 12
_DEVICE_TYPES = ("cuda",)      # the CPU test-suite adds "cpu" when it swaps in the emulation library


class Plan:
    def __init__(self, n: int, dtype: torch.dtype, device: torch.device):
        lib = _lib.load()
        self.n, self.dtype, self.device = n, dtype, device
        self.handle = ctypes.c_void_p()
        _lib.check(lib.snfft_plan_create(n, _DTYPES[dtype], device.index or 0, ctypes.byref(self.handle)),
                   "snfft_plan_create")
        npart = ctypes.c_int()
        parts = ctypes.POINTER(ctypes.c_int32)()
        dims = ctypes.POINTER(ctypes.c_int32)()
        offs = ctypes.POINTER(ctypes.c_int64)()
        _lib.check(lib.snfft_plan_info(self.handle, None, None, ctypes.byref(npart), ctypes.byref(parts),
                                       ctypes.byref(dims), ctypes.byref(offs)), "snfft_plan_info")
        self.partitions = []
        for r in range(npart.value):
            row = [parts[r * n + i] for i in range(n)]
            self.partitions.append(tuple(v for v in row if v > 0))
        self.dims = [dims[r] for r in range(npart.value)]
        self.offsets = [offs[r] for r in range(npart.value + 1)]
        self.index = {p: i for i, p in enumerate(self.partitions)}
        self.nfact = math.factorial(n)
        assert self.offsets[-1] == self.nfact

    def workspace(self, batch: int) -> torch.Tensor:
        nbytes = _lib.load().snfft_workspace_bytes(self.handle, batch)
        return torch.empty(max(nbytes, 1), dtype=torch.uint8, device=self.device)

    def __del__(self):
        try:
            if self.handle:
                _lib.load().snfft_plan_destroy(self.handle)
        except Exception:
            pass


def get_plan(n: int, dtype: torch.dtype, device: torch.device) -> Plan:
    if not isinstance(n, int) or n < 2 or n > MAX_N:
        raise ValueError(f"n must be an integer in [2, {MAX_N}], got {n!r}")
    if dtype not in _DTYPES:
        raise ValueError(f"only float32 and float64 are supported, got {dtype}")
    if device.type not in _DEVICE_TYPES:
        raise RuntimeError("algebraist_b200 runs on CUDA devices only (no CPU fallback)")
    if device.type == "cuda" and device.index is None:
        device = torch.device("cuda", torch.cuda.current_device())
    key = (n, dtype, device.type, device.index)
    with _cache_lock:
        plan = _cache.get(key)
        if plan is None:
            plan = Plan(n, dtype, device)
            _cache[key] = plan
        return plan
