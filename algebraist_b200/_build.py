"""Build the C-ABI library `libsnfft_b200.so` in-tree with nvcc for sm_100a.

    python -m algebraist_b200._build

The .so is git-ignored but travels to the GPU box with the repository
snapshot.  There is no CPU build of the product: if nvcc is missing this
raises."""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = [os.path.join(HERE, "csrc", "snfft.cu"), os.path.join(HERE, "csrc", "plan.cpp")]
DEPS = SRC + [os.path.join(HERE, "csrc", f) for f in ("kernels.cuh", "plan.h", "level_tb.cuh", "bot_gen.cuh", "base_gen.cuh", "tb_host.inl")] + [
    os.path.join(HERE, "..", "include", "snfft.h")]
OUT = os.path.join(HERE, "libsnfft_b200.so")

NVCC_FLAGS = [
    "-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC", "-shared",
]


def needs_build() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(d) > t for d in DEPS)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return OUT
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: the sm_100a library cannot be built")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", OUT] + SRC
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        sys.stderr.write(res.stderr)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
