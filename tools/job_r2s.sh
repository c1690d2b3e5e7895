# A/B of the F2 (two instances per thread, FFMA2) instantiations of the column / Z kernels against the scalar threads
TAG=$1
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-sub-records > gpurun_out/bench_${TAG}_v2.json 2> gpurun_out/bench_${TAG}_v2.err; echo "bench v2 rc=$?"
python tools/show_bench.py gpurun_out/bench_${TAG}_v2.json 2>/dev/null | head -30
SNFFT_COL_V2=0 SNFFT_TZ_ZV2=0 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-sub-records > gpurun_out/bench_${TAG}_v1.json 2> gpurun_out/bench_${TAG}_v1.err; echo "bench v1 rc=$?"
python tools/show_bench.py gpurun_out/bench_${TAG}_v1.json 2>/dev/null | head -30
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_gpu_$TAG.log
