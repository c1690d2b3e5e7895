python -m pytest tests -m gpu -x -q -k "families or golden or properties" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_gpu.log
SNFFT_TB_INV=7,8 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ia1.json 2> gpurun_out/bench_ia1.err; echo "bench rc=$?"
SNFFT_TB_INV=7,8 SNFFT_TB_INV_ALL=0 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ia0.json 2> gpurun_out/bench_ia0.err; echo "bench rc=$?"
SNFFT_TB_INV=7,8 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --dtype f64 > gpurun_out/bench_ia1_64.json 2> gpurun_out/bench_ia1_64.err; echo "bench rc=$?"
