python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r01n.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_r01n.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r01n.log 2>&1; echo "smoke rc=$?"
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r01n.json 2> gpurun_out/bench_r01n.err; echo "bench rc=$?"
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-chunk 24 > gpurun_out/bench_r01n_c24.json 2> gpurun_out/bench_r01n_c24.err; echo "bench rc=$?"
