# full measurement pass of the current build: GPU tests, smoke, bench lines, launch list, ncu --set full of every kernel
TAG=$1
COMMIT=$2
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_gpu_$TAG.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke rc=$?"
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
python tools/show_bench.py gpurun_out/bench_$TAG.json 2>/dev/null | head -30
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_${TAG}_ref.json 2> gpurun_out/bench_${TAG}_ref.err; echo "ref rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-sub-records > gpurun_out/ncu_launch_$TAG.log 2>&1; echo "launches rc=$?"
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:z_kernel|t_kernel|col_kernel|s6_kernel|gather_vec' -f -o /tmp/prof_$TAG python tools/prof_step.py --n 10 --batch 128 > gpurun_out/ncu_full_$TAG.log 2>&1; echo "full rc=$?"
python tools/ncu_summary.py /tmp/prof_$TAG.ncu-rep gpurun_out/ncu_kernels_$TAG.csv > gpurun_out/ncu_kernels_$TAG.txt; echo "summary rc=$?"
SNFFT_COMMIT=$COMMIT python tools/ncu_traffic.py gpurun_out/ncu_kernels_$TAG.csv 128 f32 gpurun_out/${TAG}_traffic.json > /dev/null
python tools/ncu_hot.py /tmp/prof_$TAG.ncu-rep 12 > gpurun_out/ncu_hot_$TAG.txt 2>&1
cut -d, -f1-6 gpurun_out/ncu_kernels_$TAG.csv
ls -la gpurun_out/ | tail -12
