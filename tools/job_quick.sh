python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3b.json 2> gpurun_out/bench_c3b.err; echo "bench rc=$?"
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --dtype f64 > gpurun_out/bench_c3b64.json 2> gpurun_out/bench_c3b64.err; echo "bench rc=$?"
