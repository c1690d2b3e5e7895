TAG=$1
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_gpu_$TAG.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke rc=$?"
python tools/pcie_bw.py > gpurun_out/pcie_$TAG.json 2>&1
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
python bench.py --steps 5 --warmup 3 --dtype f64 --no-cpu-baseline > gpurun_out/bench_${TAG}_f64.json 2> gpurun_out/bench_${TAG}_f64.err; echo "bench64 rc=$?"
python bench.py --steps 5 --warmup 3 --n 9 --batch 1024 --no-cpu-baseline > gpurun_out/bench_${TAG}_s9.json 2> gpurun_out/bench_${TAG}_s9.err; echo "bench s9 rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_${TAG}_ref.json 2> gpurun_out/bench_${TAG}_ref.err; echo "ref rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1; echo "launches rc=$?"
ncu --set full --clock-control none --kernel-name-base demangled -k 'regex:level_kernel|tb_kernel|col_kernel' -f -o /tmp/prof_final python tools/prof_step.py --n 10 --batch 128 > gpurun_out/ncu_final.log 2>&1; echo "full rc=$?"
python tools/ncu_summary.py /tmp/prof_final.ncu-rep gpurun_out/ncu_levels_$TAG.csv > /dev/null
python tools/ncu_traffic.py gpurun_out/ncu_levels_$TAG.csv 128 f32 gpurun_out/traffic_$TAG.json
python bench.py --workload coset --steps 5 --warmup 3 > gpurun_out/bench_${TAG}_coset_n1.json 2> gpurun_out/bench_${TAG}_coset_n1.err; echo "coset rc=$?"
