python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_f1.json 2> gpurun_out/bench_f1.err; echo "bench rc=$?"
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --dtype f64 > gpurun_out/bench_f1_64.json 2> gpurun_out/bench_f1_64.err; echo "bench rc=$?"
SNFFT_TB_FWD= SNFFT_TB_INV= python bench.py --steps 3 --warmup 3 --no-cpu-baseline --dtype f64 > gpurun_out/bench_f0_64.json 2> gpurun_out/bench_f0_64.err; echo "bench rc=$?"
