run() { tag=$1; shift; env "$@" python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_x_$tag.json 2> gpurun_out/bench_x_$tag.err; echo "bench $tag rc=$?"; }
run a SNFFT_COL_CHUNK_9=2 SNFFT_COL_CHUNK_8=8 SNFFT_COL_CHUNK_7=37
run b SNFFT_COL_CHUNK_9=8 SNFFT_COL_CHUNK_8=32 SNFFT_COL_CHUNK_7=148
run c SNFFT_COL_CHUNK_9=16 SNFFT_COL_CHUNK_8=64 SNFFT_COL_CHUNK_7=296 SNFFT_COL_CHUNK_6=296
run d SNFFT_COL_CHUNK_8=24 SNFFT_COL_CHUNK_6=74
