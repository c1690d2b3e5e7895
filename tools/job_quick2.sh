TAG=$1
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-sub-records > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
python tools/show_bench.py gpurun_out/bench_$TAG.json 2>/dev/null | head -24
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-sub-records --dtype f64 > gpurun_out/bench_${TAG}_f64.json 2> gpurun_out/bench_${TAG}_f64.err; echo "bench rc=$?"
python tools/show_bench.py gpurun_out/bench_${TAG}_f64.json 2>/dev/null | head -12
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
