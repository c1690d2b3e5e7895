for ch in 18 37 74 296; do
SNFFT_COL_CHUNK_9=$ch python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c9_$ch.json 2> gpurun_out/bench_c9_$ch.err; echo "bench $ch rc=$?"
done
SNFFT_COL_CHUNK_8=74 SNFFT_COL_CHUNK_7=74 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c87_74.json 2> gpurun_out/bench_c87_74.err; echo "bench rc=$?"
SNFFT_COL_CHUNK_8=100 SNFFT_COL_CHUNK_7=296 SNFFT_COL_CHUNK_6=592 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c876.json 2> gpurun_out/bench_c876.err; echo "bench rc=$?"
