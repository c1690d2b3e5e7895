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
        self._subplans = {}
        self._sub_lock = threading.Lock()

    def subplan(self, keep_indices):
        """Handle of the subset plan (snfft_subplan_create) for the partitions with these indices, cached per plan."""
        key = tuple(sorted(set(int(i) for i in keep_indices)))
        if not key or key[0] < 0 or key[-1] >= len(self.partitions):
            raise ValueError("bad partition subset")
        with self._sub_lock:
            h = self._subplans.get(key)
            if h is None:
                mask = (ctypes.c_int32 * len(self.partitions))(*[1 if i in set(key) else 0 for i in range(len(self.partitions))])
                h = ctypes.c_void_p()
                _lib.check(_lib.load().snfft_subplan_create(self.handle, mask, ctypes.byref(h)), "snfft_subplan_create")
                self._subplans[key] = h
            return h

    def subplan_info(self, handle):
        k = ctypes.c_int()
        frac = ctypes.c_double()
        _lib.check(_lib.load().snfft_subplan_info(handle, ctypes.byref(k), ctypes.byref(frac)), "snfft_subplan_info")
        return k.value, frac.value

    def workspace(self, batch: int) -> torch.Tensor:
        nbytes = _lib.load().snfft_workspace_bytes(self.handle, batch)
        return torch.empty(max(nbytes, 1), dtype=torch.uint8, device=self.device)

    def __del__(self):
        try:
            for h in getattr(self, "_subplans", {}).values():
                _lib.load().snfft_subplan_destroy(h)
            if self.handle:
                _lib.load().snfft_plan_destroy(self.handle)
        except Exception:
            pass


def clear_plan_cache():
    """Drop every cached plan (and its device tables).  The kernel-family choices of a plan are fixed when it is
    created (debug hooks, SNFFT_* environment switches), so call this after changing one of them."""
    with _cache_lock:
        _cache.clear()


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
