python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v.json 2> gpurun_out/bench_v.err; echo "bench rc=$?"
