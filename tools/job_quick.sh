python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r01q.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_r01q.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r01q.log 2>&1; echo "smoke rc=$?"
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r01q.json 2> gpurun_out/bench_r01q.err; echo "bench rc=$?"
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --dtype f64 > gpurun_out/bench_r01q_f64.json 2> gpurun_out/bench_r01q_f64.err; echo "bench64 rc=$?"
