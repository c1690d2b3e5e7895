TAG=$1
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_gpu_$TAG.log
python tools/bandlimited_time.py > gpurun_out/bandlimited_$TAG.json 2> gpurun_out/bandlimited_$TAG.err; echo "band rc=$?"; cat gpurun_out/bandlimited_$TAG.json
python tools/sanitize_cases.py > gpurun_out/cases_$TAG.log 2>&1; echo "cases rc=$?"; tail -2 gpurun_out/cases_$TAG.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
python tools/show_bench.py gpurun_out/bench_$TAG.json | head -6
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_${TAG}_ref.json 2> gpurun_out/bench_${TAG}_ref.err; echo "ref rc=$?"; cat gpurun_out/bench_${TAG}_ref.json | cut -c1-600
