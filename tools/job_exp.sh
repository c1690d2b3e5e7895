TAG=$1
for cw in 8 16 32; do
SNFFT_TZ_PK_CW=$cw python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-sub-records > gpurun_out/bench_${TAG}_cw$cw.json 2> gpurun_out/bench_${TAG}_cw$cw.err; echo "bench cw$cw rc=$?"
python tools/show_bench.py gpurun_out/bench_${TAG}_cw$cw.json 2>/dev/null | grep -E 'transforms|k10_c-7'
done
