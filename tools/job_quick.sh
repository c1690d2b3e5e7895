python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest_col2.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_col2.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_col2.json 2> gpurun_out/bench_col2.err; echo "bench rc=$?"
SNFFT_COL_CHUNK=74 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_col2_ch74.json 2> gpurun_out/bench_col2_ch74.err; echo "bench rc=$?"
SNFFT_COL_CHUNK=296 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_col2_ch296.json 2> gpurun_out/bench_col2_ch296.err; echo "bench rc=$?"
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --dtype f64 > gpurun_out/bench_col2_64.json 2> gpurun_out/bench_col2_64.err; echo "bench64 rc=$?"
