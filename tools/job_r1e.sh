set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r1e.json 2> gpurun_out/bench_r1e.err; echo "bench rc=$?"
python tools/prof_step.py --n 10 --batch 32 > gpurun_out/prof_plain.log 2>&1; echo "plain rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1e.csv python tools/prof_step.py --n 10 --batch 32 > gpurun_out/ncu_l.log 2>&1; echo "launches rc=$?"
ncu --set full --clock-control none --import-source on -k regex:level_kernel -f -o gpurun_out/prof_levels_r1e python tools/prof_step.py --n 10 --batch 32 > gpurun_out/ncu_f.log 2>&1; echo "full rc=$?"
