python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_gpu.log
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r1g.json 2> gpurun_out/bench_r1g.err; echo "bench rc=$?"
python bench.py --steps 3 --warmup 3 --dtype f64 --no-cpu-baseline > gpurun_out/bench_r1g_f64.json 2> gpurun_out/bench_r1g_f64.err; echo "bench rc=$?"
