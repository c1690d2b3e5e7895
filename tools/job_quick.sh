python -m pytest tests -m gpu -x -q -k "coset or n11" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_gpu.log
python bench.py --workload coset --steps 3 --warmup 3 > gpurun_out/bench_coset_n1.json 2> gpurun_out/bench_coset_n1.err; echo "coset1 rc=$?"
tail -3 gpurun_out/bench_coset_n1.err
cat gpurun_out/bench_coset_n1.json
