# packed T kernels: warp-tile width 16 / 8 / 32 (tuning build), level-8 inverse with one coset per task for d = 35
TAG=$1
for cw in 16 8 32; do
SNFFT_TZ_PK_CW=$cw python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-sub-records > gpurun_out/bench_${TAG}_cw$cw.json 2> gpurun_out/bench_${TAG}_cw$cw.err; echo "bench cw$cw rc=$?"
python tools/show_bench.py gpurun_out/bench_${TAG}_cw$cw.json 2>/dev/null | grep -E 'transforms|k10_c-6|k8_c-3'
done
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "oracle_large or properties" > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_gpu_$TAG.log
