TAG=$1
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-sub-records > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; echo "bench rc=$?"
python tools/show_bench.py gpurun_out/bench_${TAG}.json 2>/dev/null | head -20
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_gpu_$TAG.log
