"""The fast paths are re-arrangements of the same arithmetic, not different arithmetic: with any of them switched
off (scalar threads instead of packed fp32 pairs, separate pack / unpack instead of the packed-layout T kernel,
levels 2..5 and 6 as two kernels instead of one) the results are BIT-IDENTICAL.  The switches are read once per
process, so every configuration runs in its own interpreter."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

CONFIGS = {
    "default": {},
    "scalar_threads": {"SNFFT_COL_V2": "0", "SNFFT_TZ_ZV2": "0"},
    "separate_pack": {"SNFFT_TZ_PACK": "0"},
    "separate_levels_5_6": {"SNFFT_S6": "0"},
    "all_off": {"SNFFT_COL_V2": "0", "SNFFT_TZ_ZV2": "0", "SNFFT_TZ_PACK": "0", "SNFFT_S6": "0"},
}


def _run(env_extra):
    env = dict(os.environ)
    env.update(env_extra)
    res = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "workers", "switch_worker.py")], capture_output=True,
                         text=True, timeout=900, env=env)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    cases = [ln for ln in res.stdout.splitlines() if ln.startswith("CASE ")]
    assert len(cases) == 4, res.stdout
    return cases


@pytest.mark.gpu
def test_kernel_switches_are_bit_identical():
    ref = _run(CONFIGS["default"])
    for ln in ref:
        assert float(ln.split("err=")[1]) < 1e-5, ln
    for name, env in CONFIGS.items():
        if name == "default":
            continue
        got = _run(env)
        assert [ln.split(" err=")[0] for ln in got] == [ln.split(" err=")[0] for ln in ref], (name, got, ref)
