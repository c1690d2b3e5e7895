# fused levels-2..6 kernel (s6_kernels.cuh) against the s5 + level-6 column kernels
TAG=$1
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-sub-records > gpurun_out/bench_${TAG}_s6.json 2> gpurun_out/bench_${TAG}_s6.err; echo "bench s6 rc=$?"
python tools/show_bench.py gpurun_out/bench_${TAG}_s6.json 2>/dev/null | head -30
SNFFT_S6=0 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-sub-records > gpurun_out/bench_${TAG}_nos6.json 2> gpurun_out/bench_${TAG}_nos6.err; echo "bench nos6 rc=$?"
python tools/show_bench.py gpurun_out/bench_${TAG}_nos6.json 2>/dev/null | head -3
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_gpu_$TAG.log
