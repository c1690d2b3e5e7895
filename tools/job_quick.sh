for cfg in "4" "16" "32" "64"; do set -- $cfg
SNFFT_COL_CHUNK_9=4 SNFFT_COL_CHUNK_8=$1 SNFFT_COL_CHUNK_7=$1 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v_$1.json 2> gpurun_out/bench_v_$1.err; echo "bench $cfg rc=$?"
done
for cfg in "4 8" "8 24" "16 48"; do set -- $cfg
SNFFT_COL_CHUNK_9=$1 SNFFT_COL_CHUNK_8=$2 SNFFT_COL_CHUNK_7=$2 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --dtype f64 > gpurun_out/bench_v64_$1.json 2> gpurun_out/bench_v64_$1.err; echo "bench rc=$?"
done
