TAG=$1
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-sub-records > gpurun_out/bench_${TAG}_a.json 2> gpurun_out/bench_${TAG}_a.err; echo "bench a rc=$?"
python tools/show_bench.py gpurun_out/bench_${TAG}_a.json 2>/dev/null | grep -E 'transforms|k10_c-7|k6_c-8'
SNFFT_S6_V2=0 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-sub-records > gpurun_out/bench_${TAG}_b.json 2> gpurun_out/bench_${TAG}_b.err; echo "bench b rc=$?"
python tools/show_bench.py gpurun_out/bench_${TAG}_b.json 2>/dev/null | grep -E 'transforms|k10_c-7|k6_c-8'
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden or oracle_large or ragged" > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_gpu_$TAG.log
