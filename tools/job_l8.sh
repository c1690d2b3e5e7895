TAG=$1
shift
for ch in "$@"; do
SNFFT_COL_CHUNK2_8=$ch python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-sub-records > gpurun_out/bench_${TAG}_ch$ch.json 2> gpurun_out/bench_${TAG}_ch$ch.err; echo "bench ch$ch rc=$?"
python tools/show_bench.py gpurun_out/bench_${TAG}_ch$ch.json 2>/dev/null | grep -E 'transforms|k8_c-3'
done
