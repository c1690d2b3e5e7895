python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r01p.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_r01p.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r01p.json 2> gpurun_out/bench_r01p.err; echo "bench rc=$?"
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --dtype f64 > gpurun_out/bench_r01p_f64.json 2> gpurun_out/bench_r01p_f64.err; echo "bench64 rc=$?"
python bench.py --steps 5 --warmup 3 --n 9 --batch 1024 --no-cpu-baseline > gpurun_out/bench_r01p_s9.json 2> gpurun_out/bench_r01p_s9.err; echo "bench s9 rc=$?"
