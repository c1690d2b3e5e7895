python bench.py --steps 3 --warmup 3 --no-cpu-baseline --dtype f64 > gpurun_out/bench_e1.json 2> gpurun_out/bench_e1.err; echo "bench rc=$?"
SNFFT_TB_FWD= SNFFT_TB_INV= python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_e4.json 2> gpurun_out/bench_e4.err; echo "bench rc=$?"
SNFFT_TB_FWD= SNFFT_TB_INV= python bench.py --steps 3 --warmup 3 --no-cpu-baseline --dtype f64 > gpurun_out/bench_e4_64.json 2> gpurun_out/bench_e4_64.err; echo "bench rc=$?"
