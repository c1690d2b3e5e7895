python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r01r.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_r01r.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r01r.log 2>&1; echo "smoke rc=$?"
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r01r.json 2> gpurun_out/bench_r01r.err; echo "bench rc=$?"
