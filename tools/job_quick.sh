SNFFT_TB_FWD= SNFFT_TB_INV= python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_e5.json 2> gpurun_out/bench_e5.err; echo "bench rc=$?"
