SNFFT_TB_FWD= SNFFT_TB_INV= python bench.py --steps 3 --warmup 3 --no-cpu-baseline --dtype f64 > gpurun_out/bench_d0.json 2> gpurun_out/bench_d0.err; echo "bench rc=$?"
SNFFT_TB_FWD=7,8,9,10 SNFFT_TB_INV=7,8,9,10 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --dtype f64 > gpurun_out/bench_d1.json 2> gpurun_out/bench_d1.err; echo "bench rc=$?"
SNFFT_TB_WAVES=0 SNFFT_TB_FWD=7,8,9,10 SNFFT_TB_INV=7,8,9,10 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --dtype f64 > gpurun_out/bench_d2.json 2> gpurun_out/bench_d2.err; echo "bench rc=$?"
