# A/B of the packed-layout T kernel (pack / unpack fused into the top level's T kernel) against the two-pass route
TAG=$1
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-sub-records > gpurun_out/bench_${TAG}_pk.json 2> gpurun_out/bench_${TAG}_pk.err; echo "bench pk rc=$?"
python tools/show_bench.py gpurun_out/bench_${TAG}_pk.json 2>/dev/null | head -30
SNFFT_TZ_PACK=0 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-sub-records > gpurun_out/bench_${TAG}_nopk.json 2> gpurun_out/bench_${TAG}_nopk.err; echo "bench nopk rc=$?"
python tools/show_bench.py gpurun_out/bench_${TAG}_nopk.json 2>/dev/null | head -6
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_gpu_$TAG.log
