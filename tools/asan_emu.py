"""AddressSanitizer / UBSan run of the kernel logic on the CPU (the pool's compute-sanitizer is closed, see
profiles/r02g_compute_sanitizer_closed.log): the CUDA sources are compiled by g++ against tests/emu/cuda_emu.h with
-fsanitize=address,undefined, so every global- and shared-memory access of every emulated thread is bounds checked
(device buffers are host allocations, dynamic shared memory is a heap vector), and the cases of the emulation
test-suite run on that build.

    python tools/asan_emu.py            # builds tests/emu/_build_asan/libsnfft_emu_asan.so, runs the cases, prints a summary
"""
import ctypes
import math
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
EMU_DIR = os.path.join(ROOT, "tests", "emu")
BDIR = os.path.join(EMU_DIR, "_build_asan")
SO = os.path.join(BDIR, "libsnfft_emu_asan.so")
SAN = ["-fsanitize=address,undefined", "-fno-omit-frame-pointer", "-g1"]


def build():
    from algebraist_b200 import _build as PB
    os.makedirs(BDIR, exist_ok=True)
    PB.generate_col_headers()
    units = [(os.path.join(BDIR, os.path.basename(o)), s, e, d) for o, s, e, d in PB.units()]
    todo = [u for u in units if PB.stale(u[0], u[3])]

    def one(u):
        obj, src, extra, _ = u
        subprocess.run(["g++", "-std=c++17", "-O1", "-x", "c++", "-DSNFFT_HOST_EMU", "-I" + EMU_DIR, "-fPIC"] + SAN + extra +
                       ["-c", "-o", obj, src], check=True)

    with ThreadPoolExecutor(max_workers=os.cpu_count() or 1) as ex:
        list(ex.map(one, todo))
    if todo or not os.path.exists(SO):
        subprocess.run(["g++", "-shared"] + SAN + ["-o", SO] + [u[0] for u in units], check=True)


def cases():
    import numpy as np
    from conftest import CAbi, unpack, rel_err
    from algebraist_b200 import _lib
    from oracle import sn_oracle as O
    emu = _lib.bind(ctypes.CDLL(SO))
    abi = CAbi(emu)
    done = []

    def run(n, B, dt, tag):
        rng = np.random.default_rng(n * 10 + B)
        f = rng.standard_normal((B, math.factorial(n))).astype(dt)
        plan = abi.plan(n, dt)
        try:
            packed = abi.forward(plan, f)
            ref = O.sn_fft_packed_c(f.astype(np.float64), n)
            ft = unpack(packed, n, B)
            tol = 1e-11 if dt == np.float64 else 2e-5
            assert max(rel_err(ft[l], ref[l]) for l in ref) < tol, (tag, n, B)
            assert rel_err(abi.inverse(plan, packed, B), f) < tol
        finally:
            emu.snfft_plan_destroy(plan)
        done.append(f"{tag}: n={n} B={B} {np.dtype(dt).name}")
        print(done[-1], flush=True)

    for n, B, dt in ((6, 5, np.float32), (7, 4, np.float32), (7, 3, np.float64), (8, 2, np.float64), (8, 4, np.float32),
                     (9, 1, np.float64), (9, 4, np.float32), (9, 3, np.float32),
                     # (6, 4): levels 2..6 kernel + pack; (9, 2) fp64 / (9, 4) fp32 above: T kernel on the packed layout,
                     # F2 column / Z kernels (fp32)
                     (6, 4, np.float32), (9, 2, np.float64)):
        run(n, B, dt, "auto")
    for mode in (0,):
        emu.snfft_debug_tb_mode(mode)
        run(7, 5, np.float64, f"tb_mode={mode}")
        run(8, 1, np.float32, f"tb_mode={mode}")
    emu.snfft_debug_tb_mode(-1)
    for base in (1,):
        emu.snfft_debug_disable_base(base)
        run(7, 3, np.float64, f"base={base}")
    emu.snfft_debug_disable_base(0)
    emu.snfft_debug_scalar_gather(1)
    run(7, 4, np.float32, "scalar gather/pack")
    emu.snfft_debug_scalar_gather(0)
    print("ASAN_EMU_CASES_OK", len(done))


if __name__ == "__main__":
    if "--run" in sys.argv:
        cases()
    else:
        build()
        asan = subprocess.run(["g++", "-print-file-name=libasan.so"], capture_output=True, text=True).stdout.strip()
        env = dict(os.environ, LD_PRELOAD=asan, ASAN_OPTIONS="detect_leaks=0:detect_stack_use_after_return=0",
                   UBSAN_OPTIONS="print_stacktrace=1:halt_on_error=1")
        sys.exit(subprocess.run([sys.executable, os.path.abspath(__file__), "--run"], env=env).returncode)
