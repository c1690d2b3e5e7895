"""ctypes binding of include/snfft.h (the C ABI of the CUDA library).

The library is the product: there is no CPU or PyTorch fallback.  Loading
fails loudly if `libsnfft_b200.so` has not been built
(`python -m algebraist_b200._build`)."""
import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsnfft_b200.so")

_lock = threading.Lock()
_lib = None

c_vp = ctypes.c_void_p
c_i64 = ctypes.c_int64
c_int = ctypes.c_int

class ProfileRecord(ctypes.Structure):
    _fields_ = [("kind", ctypes.c_int32), ("level", ctypes.c_int32), ("kernel_class", ctypes.c_int32),
                ("launches", ctypes.c_int32), ("elems_per_instance", ctypes.c_int64),
                ("total_ms", ctypes.c_double)]


# name -> (restype, argtypes); mirrors include/snfft.h one to one
SIGNATURES = {
    "snfft_build_info": (ctypes.c_char_p, []),
    "snfft_last_error": (ctypes.c_char_p, []),
    "snfft_plan_create": (c_int, [c_int, c_int, c_int, ctypes.POINTER(c_vp)]),
    "snfft_plan_destroy": (None, [c_vp]),
    "snfft_plan_info": (c_int, [c_vp, ctypes.POINTER(c_int), ctypes.POINTER(c_int), ctypes.POINTER(c_int),
                                ctypes.POINTER(ctypes.POINTER(ctypes.c_int32)),
                                ctypes.POINTER(ctypes.POINTER(ctypes.c_int32)),
                                ctypes.POINTER(ctypes.POINTER(ctypes.c_int64))]),
    "snfft_plan_level_info": (c_int, [c_vp, c_int, ctypes.POINTER(c_int),
                                      ctypes.POINTER(ctypes.POINTER(ctypes.c_int32)),
                                      ctypes.POINTER(ctypes.POINTER(ctypes.c_int32))]),
    "snfft_plan_yor_table": (c_int, [c_vp, c_int, c_int, c_int, ctypes.POINTER(c_int),
                                     ctypes.POINTER(ctypes.POINTER(ctypes.c_int32)),
                                     ctypes.POINTER(ctypes.POINTER(ctypes.c_double)),
                                     ctypes.POINTER(ctypes.POINTER(ctypes.c_double))]),
    "snfft_workspace_bytes": (ctypes.c_size_t, [c_vp, c_i64]),
    "snfft_forward": (c_int, [c_vp, c_vp, c_i64, c_vp, c_vp, ctypes.c_size_t, c_vp]),
    "snfft_inverse": (c_int, [c_vp, c_vp, c_i64, c_vp, c_vp, ctypes.c_size_t, c_vp]),
    "snfft_forward_cosets_workspace_bytes": (ctypes.c_size_t, [c_vp, c_vp, c_i64, c_int]),
    "snfft_forward_cosets2_workspace_bytes": (ctypes.c_size_t, [c_vp, c_vp, c_i64, c_int]),
    "snfft_forward_cosets2": (c_int, [c_vp, c_vp, c_vp, c_i64, ctypes.POINTER(c_int), ctypes.POINTER(c_int), c_int, c_vp, c_vp,
                                      ctypes.c_size_t, c_vp]),
    "snfft_forward_cosets": (c_int, [c_vp, c_vp, c_vp, c_i64, ctypes.POINTER(c_int), c_int, c_vp, c_vp, ctypes.c_size_t,
                                     c_vp]),
    "snfft_subplan_create": (c_int, [c_vp, ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(c_vp)]),
    "snfft_subplan_destroy": (None, [c_vp]),
    "snfft_subplan_info": (c_int, [c_vp, ctypes.POINTER(c_int), ctypes.POINTER(ctypes.c_double)]),
    "snfft_forward_subset": (c_int, [c_vp, c_vp, c_vp, c_i64, c_vp, c_vp, ctypes.c_size_t, c_vp]),
    "snfft_inverse_subset": (c_int, [c_vp, c_vp, c_vp, c_i64, c_vp, c_vp, ctypes.c_size_t, c_vp]),
    "snfft_power": (c_int, [c_vp, c_vp, c_i64, c_vp, c_vp]),
    "snfft_block_matmul": (c_int, [c_vp, c_vp, c_vp, c_vp, c_i64, c_vp]),
    "snfft_forward_host": (c_int, [c_vp, c_vp, c_i64, c_vp, c_vp]),
    "snfft_inverse_host": (c_int, [c_vp, c_vp, c_i64, c_vp, c_vp]),
    "snfft_perm_rank": (c_int, [c_int, c_vp, c_i64, c_vp, c_vp]),
    "snfft_perm_unrank": (c_int, [c_int, c_vp, c_i64, c_vp, c_vp]),
    "snfft_coset_index": (c_int, [c_int, c_int, c_vp, c_vp]),
    "snfft_chain_index": (c_int, [c_int, c_vp, c_vp]),
    "snfft_dense_rep": (c_int, [c_vp, c_int, c_vp, c_vp]),
    "snfft_rho": (c_int, [c_vp, c_int, ctypes.POINTER(c_int), c_vp, c_vp]),
    "snfft_translate": (c_int, [c_int, c_int, c_vp, c_i64, ctypes.POINTER(c_int), c_int, c_vp, c_vp]),
    "snfft_profile_begin": (c_int, [c_vp]),
    "snfft_profile_end": (c_int, [c_vp, ctypes.POINTER(ProfileRecord), c_int, ctypes.POINTER(c_int)]),
    "snfft_launch_count": (c_i64, []),
    "snfft_debug_force_class": (None, [c_int]),
    "snfft_debug_disable_base": (None, [c_int]),
    "snfft_debug_tb_mode": (None, [c_int]),
    "snfft_debug_scalar_gather": (None, [c_int]),
}


def bind(lib):
    """Attach the header's signatures to a loaded library."""
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    return lib


def load():
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise RuntimeError(
                    f"{LIB_PATH} is missing: build the CUDA library with "
                    "`python -m algebraist_b200._build` (there is no CPU fallback)")
            _lib = bind(ctypes.CDLL(LIB_PATH))
        return _lib


def check(rc: int, what: str):
    if rc == 0:
        return
    msg = load().snfft_last_error().decode(errors="replace")
    if rc in (-1,):
        raise ValueError(f"{what}: {msg}")
    if rc == -2:
        raise MemoryError(f"{what}: {msg}")
    raise RuntimeError(f"{what} failed ({rc}): {msg}")
