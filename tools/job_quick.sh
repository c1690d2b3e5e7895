python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "ragged or config2 or n12" > gpurun_out/pytest_edge.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_edge.log
python bench.py --steps 5 --warmup 3 --n 8 --batch 4096 --no-cpu-baseline > gpurun_out/bench_s8_b4096.json 2> gpurun_out/bench_s8_b4096.err; echo "bench s8 rc=$?"
