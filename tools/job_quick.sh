for ch in 1 16 148 592 2368; do
SNFFT_COL_CHUNK=$ch python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ch$ch.json 2> gpurun_out/bench_ch$ch.err; echo "bench $ch rc=$?"
done
