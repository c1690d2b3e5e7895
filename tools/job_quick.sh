python -m pytest tests -m gpu -x -q -k "families or golden" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_gpu.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ws1.json 2> gpurun_out/bench_ws1.err; echo "bench rc=$?"
SNFFT_TB_FWD=7,8,9,10 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ws2.json 2> gpurun_out/bench_ws2.err; echo "bench rc=$?"
SNFFT_TB_WS=0 SNFFT_TB_FWD=7,8,9,10 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ws3.json 2> gpurun_out/bench_ws3.err; echo "bench rc=$?"
