python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest_col3.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_col3.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_col3.json 2> gpurun_out/bench_col3.err; echo "bench rc=$?"
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --dtype f64 > gpurun_out/bench_col3_64.json 2> gpurun_out/bench_col3_64.err; echo "bench64 rc=$?"
python bench.py --steps 5 --warmup 3 --n 9 --batch 1024 --no-cpu-baseline > gpurun_out/bench_col3_s9.json 2> gpurun_out/bench_col3_s9.err; echo "bench s9 rc=$?"
