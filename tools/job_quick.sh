python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_gpu.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg_auto.json 2> gpurun_out/bench_cfg_auto.err; echo "bench rc=$?"
SNFFT_TB_FWD= SNFFT_TB_INV= python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg_old.json 2> gpurun_out/bench_cfg_old.err; echo "bench rc=$?"
SNFFT_TB_FWD=7,8,9,10 SNFFT_TB_INV=7,8,9,10 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg_tb.json 2> gpurun_out/bench_cfg_tb.err; echo "bench rc=$?"
SNFFT_TB_FWD=7,8,9,10 SNFFT_TB_INV=7,8,9,10 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --dtype f64 > gpurun_out/bench_cfg_tb64.json 2> gpurun_out/bench_cfg_tb64.err; echo "bench rc=$?"
SNFFT_TB_FWD= SNFFT_TB_INV= python bench.py --steps 3 --warmup 3 --no-cpu-baseline --dtype f64 > gpurun_out/bench_cfg_old64.json 2> gpurun_out/bench_cfg_old64.err; echo "bench rc=$?"
