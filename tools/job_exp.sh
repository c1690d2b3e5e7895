TAG=$1
for cw in 8 16 32; do
SNFFT_TZ_PK_CW=$cw python bench.py --n 9 --batch 1024 --steps 10 --warmup 3 --no-cpu-baseline --no-sub-records > gpurun_out/bench_${TAG}_s9_cw$cw.json 2> gpurun_out/bench_${TAG}_s9_cw$cw.err; echo "bench cw$cw rc=$?"
python tools/show_bench.py gpurun_out/bench_${TAG}_s9_cw$cw.json 2>/dev/null | grep -E 'transforms|k9_c-7'
done
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "full_size" > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_gpu_$TAG.log
