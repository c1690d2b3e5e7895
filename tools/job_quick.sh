SNFFT_TB_FWD=9 SNFFT_TB_INV= python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c0.json 2> gpurun_out/bench_c0.err; echo "bench rc=$?"
