TAG=$1
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
python tools/show_bench.py gpurun_out/bench_$TAG.json
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_gpu_$TAG.log
python bench.py --steps 5 --warmup 3 --dtype f64 --no-cpu-baseline > gpurun_out/bench_${TAG}_f64.json 2> gpurun_out/bench_${TAG}_f64.err; echo "bench64 rc=$?"
python tools/show_bench.py gpurun_out/bench_${TAG}_f64.json
python bench.py --steps 5 --warmup 3 --n 9 --batch 1024 --no-cpu-baseline > gpurun_out/bench_${TAG}_s9.json 2> gpurun_out/bench_${TAG}_s9.err; echo "bench s9 rc=$?"
python tools/show_bench.py gpurun_out/bench_${TAG}_s9.json
