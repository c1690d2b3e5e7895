"""Build the C-ABI library `libsnfft_b200.so` in-tree with nvcc for sm_100a.

    python -m algebraist_b200._build [--force] [-v]

The generated column kernels (csrc/col_tu.cu) are compiled as one translation unit per
(level, dtype, direction), in parallel with the main one; objects are cached under `_obj/`.
The .so is git-ignored but travels to the GPU box with the repository snapshot.  There is no
CPU build of the product: if nvcc is missing this raises."""
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "_obj")
OUT = os.path.join(HERE, "libsnfft_b200.so")
HEADER = os.path.join(HERE, "..", "include", "snfft.h")

MAIN_DEPS = [os.path.join(CSRC, f) for f in ("snfft.cu", "kernels.cuh", "plan.h", "base_gen.cuh")] + [HEADER]
S6_DEPS = [os.path.join(CSRC, f) for f in ("s6_tu.cu", "s6_kernels.cuh", "s6_gen.cuh", "base_gen.cuh", "vec2.cuh")]
PLAN_DEPS = [os.path.join(CSRC, f) for f in ("plan.cpp", "plan.h")]
def col_deps(k):
    return [os.path.join(CSRC, f) for f in ("col_tu.cu", "col_kernels.cuh", f"col_gen_{k}.cuh")]



def tz_deps(kind):
    gen = f"tz_gen_z{kind}.cuh" if kind <= 8 else f"tz_gen_t{kind}.cuh"
    return [os.path.join(CSRC, f) for f in ("tz_tu.cu", "tz_kernels.cuh", "tz_gen_tab.cuh", gen)]


# top / bottom kernels (csrc/tz_tu.cu): kind 7 / 8 = Z kernel with that bottom level, 9 / 10 / 11 = T kernel of that level
TZ_COMBOS = [(kind, f64, inv) for kind in (8, 7, 11, 10, 9) for f64 in (0, 1) for inv in (0, 1)]
COL_COMBOS = [(k, f64, inv) for k in (6, 7, 8) for f64 in (0, 1) for inv in (0, 1)]
# threads per CTA of the column kernels by (level, fp64); default 128.  One large CTA per SM keeps all its warps on
# the code of ONE task (instruction fetch was the top stall at levels 8 and 9 with five 128-thread CTAs of
# different tasks per SM): forward k = 9 3.36 -> 2.30 ms, fp64 5.6 -> 3.7 ms; level 7 is faster with 128.
COL_THREADS = {(8, 0): 640, (8, 1): 384}

if os.environ.get("SNFFT_T_THREADS"):           # tuning build: CTA size of the T kernels
    TZ_T = [f"-DSNFFT_T_THREADS={int(os.environ['SNFFT_T_THREADS'])}"]
else:
    TZ_T = []
if os.environ.get("SNFFT_Z_THREADS"):           # tuning build: CTA size of the Z kernels
    TZ_T += [f"-DSNFFT_Z_THREADS={int(os.environ['SNFFT_Z_THREADS'])}"]
if os.environ.get("SNFFT_TZ_PK_TUNE"):          # tuning build: all three warp-tile widths of the packed T kernels (SNFFT_TZ_PK_CW picks)
    TZ_EXTRA = ["-DSNFFT_TZ_PK_TUNE=1"]
else:
    TZ_EXTRA = []
# threads per CTA of the F2 instantiation (fp32, two instances per thread) by level; default min(COL_THREADS, 384)
COL_THREADS_V2 = {}


# resident threads per SM the F2 instantiations are compiled for (register cap 65536 / this; default 384 = 168 registers).
# With 48 running sums per forward task the level-7 code fits 128 registers: four CTAs of 128 threads per SM
# (level 7 forward 0.67 -> 0.64 ms; the inverse tasks spill under that cap: 0.66 -> 0.67 ms)
COL_V2_REGS_THREADS = {(7, 0): 512}
Z_V2_REGS_THREADS = {}          # the Z kernels (run-time strides, segment lookup) lose under the same cap: 0.55 -> 0.57 ms


def col_v2_flags(k, f64, inv=0):
    t = os.environ.get(f"SNFFT_COL{k}_V2_THREADS") or COL_THREADS_V2.get(k)
    out = [f"-DSNFFT_COL_THREADS_V2={int(t)}"] if t and not f64 else []
    r = os.environ.get(f"SNFFT_COL{k}_V2_REGS_THREADS") or COL_V2_REGS_THREADS.get((k, inv))
    if r and not f64:
        out.append(f"-DSNFFT_COL_V2_REGS_THREADS={int(r)}")
    return out


def z_v2_flags(kind, f64, inv):
    r = os.environ.get("SNFFT_Z_V2_REGS_THREADS") or Z_V2_REGS_THREADS.get((kind, inv))
    return [f"-DSNFFT_Z_V2_REGS_THREADS={int(r)}"] if r and not f64 and kind in (7, 8) else []


NVCC_FLAGS = ["-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-Xcompiler", "-fPIC"]


def units():
    """(object, source, extra flags, dependencies); the largest translation units first."""
    u = [(os.path.join(OBJ, f"col_{k}_{'f64' if f64 else 'f32'}_{'inv' if inv else 'fwd'}_t{COL_THREADS.get((k, f64), 128)}.o"), os.path.join(CSRC, "col_tu.cu"),
          [f"-DSNFFT_COL_K={k}", f"-DSNFFT_COL_F64={f64}", f"-DSNFFT_COL_INV={inv}",
           f"-DSNFFT_COL_THREADS={COL_THREADS.get((k, f64), 128)}"] + col_v2_flags(k, f64, inv), col_deps(k))
         for k, f64, inv in sorted(COL_COMBOS, reverse=True)]
    u += [(os.path.join(OBJ, f"tz_{kind}_{'f64' if f64 else 'f32'}_{'inv' if inv else 'fwd'}.o"), os.path.join(CSRC, "tz_tu.cu"),
           [f"-DSNFFT_TZ_KIND={kind}", f"-DSNFFT_TZ_F64={f64}", f"-DSNFFT_TZ_INV={inv}"] + TZ_EXTRA + TZ_T + z_v2_flags(kind, f64, inv), tz_deps(kind))
          for kind, f64, inv in TZ_COMBOS]
    u += [(os.path.join(OBJ, f"s6_f32_{'inv' if inv else 'fwd'}.o"), os.path.join(CSRC, "s6_tu.cu"), [f"-DSNFFT_S6_INV={inv}"], S6_DEPS)
          for inv in (0, 1)]
    u.insert(0, (os.path.join(OBJ, "snfft.o"), os.path.join(CSRC, "snfft.cu"), [], MAIN_DEPS))
    u.append((os.path.join(OBJ, "plan.o"), os.path.join(CSRC, "plan.cpp"), [], PLAN_DEPS))
    return u


def generate_col_headers(force: bool = False):
    """col_gen_<k>.cuh (20 MB of straight-line code) are not kept in git: csrc/gen_col.py writes them (2 s)."""
    for script, gen_names, head_names in (
            ("gen_col.py", ("gen_col.py", "gen_base.py"),
             [f"col_gen_{k}.cuh" for k in (6, 7, 8)] + ["s6_gen.cuh"]),
            ("gen_tz.py", ("gen_tz.py", "gen_col.py", "gen_base.py"),
             ["tz_gen_z7.cuh", "tz_gen_z8.cuh", "tz_gen_tab.cuh", "tz_gen_t9.cuh", "tz_gen_t10.cuh", "tz_gen_t11.cuh"])):
        gens = [os.path.join(CSRC, f) for f in gen_names]
        heads = [os.path.join(CSRC, f) for f in head_names]
        # the generators only rewrite a header whose contents change, so an edit of a generator that leaves the
        # code alone does not recompile anything; the stamp records that the generator ran after its last edit
        stamp = os.path.join(OBJ, script + ".stamp")
        if force or stale(stamp, gens) or not all(os.path.exists(h) for h in heads):
            res = subprocess.run([sys.executable, os.path.join(CSRC, script)], capture_output=True, text=True, cwd=CSRC)
            if res.returncode != 0:
                raise RuntimeError(script + " failed:\n" + res.stdout + res.stderr)
            os.makedirs(OBJ, exist_ok=True)
            with open(stamp, "w") as fh:
                fh.write("generated\n")


def stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def needs_build() -> bool:
    """The library is up to date when it is newer than every source it was built from; the object cache under
    _obj/ only matters for an incremental rebuild (it does not travel to the GPU box: .gpurunignore)."""
    generate_col_headers()
    deps = sorted({d for _, _, _, ds in units() for d in ds})
    return stale(OUT, deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return OUT
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: the sm_100a library cannot be built")
    os.makedirs(OBJ, exist_ok=True)
    generate_col_headers()
    todo = [u for u in units() if force or stale(u[0], u[3])]

    def compile_one(u):
        obj, src, extra, _ = u
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + extra + ["-c", "-o", obj, src]
        return obj, subprocess.run(cmd, capture_output=True, text=True)

    with ThreadPoolExecutor(max_workers=max(1, min(len(todo), os.cpu_count() or 1))) as ex:
        for obj, res in ex.map(compile_one, todo):
            if res.returncode != 0:
                raise RuntimeError(f"nvcc failed for {os.path.basename(obj)}:\n" + res.stdout + res.stderr)
            if verbose:
                sys.stderr.write(res.stderr)
    res = subprocess.run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", OUT] + [u[0] for u in units()],
                         capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("link failed:\n" + res.stdout + res.stderr)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
