TAG=$1
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-sub-records > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; echo "bench rc=$?"
python tools/show_bench.py gpurun_out/bench_${TAG}.json 2>/dev/null | grep -E 'transforms|k7_c-3|c-4'
