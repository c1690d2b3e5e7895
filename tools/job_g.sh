python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-sub-records > gpurun_out/bench_r04w.json 2> gpurun_out/bench_r04w.err; echo "bench rc=$?"
python tools/show_bench.py gpurun_out/bench_r04w.json 2>/dev/null | grep -E 'transforms|gather|scatter'
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden or ragged or host" 2>&1 | tail -2
