#!/bin/bash
# Regenerate the column / top-bottom headers with the generator switches of the environment (SNFFT_GEN_*) and rebuild.
# usage: SNFFT_GEN_FWD_ROWS_7=48 bash tools/regen_build.sh
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
(cd "$ROOT/algebraist_b200/csrc" && python gen_col.py > /dev/null && python gen_tz.py > /dev/null)
cd "$ROOT" && python -m algebraist_b200._build -v > /tmp/regen_build.log 2>&1
grep -c 'Compiling entry' /tmp/regen_build.log || true
tail -1 /tmp/regen_build.log
