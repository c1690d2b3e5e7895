TAG=$1
for ch in 18 37 74 148 296; do
SNFFT_COL_CHUNK2_7=$ch SNFFT_TZ_CHUNK2=$ch python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-sub-records > gpurun_out/bench_${TAG}_ch$ch.json 2> gpurun_out/bench_${TAG}_ch$ch.err; echo "bench ch$ch rc=$?"
python tools/show_bench.py gpurun_out/bench_${TAG}_ch$ch.json 2>/dev/null | grep -E 'transforms|c-4|k7_c-3'
done
