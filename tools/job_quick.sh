(nvidia-smi topo -m; cat /sys/devices/system/node/node*/cpulist; python -c "import os; print(len(os.sched_getaffinity(0)))") > gpurun_out/topo.txt 2>&1
for i in 1 2; do
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_e2e_a$i.json 2> gpurun_out/bench_e2e_a$i.err; echo "bench rc=$?"
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-numa-bind > gpurun_out/bench_e2e_b$i.json 2> gpurun_out/bench_e2e_b$i.err; echo "bench rc=$?"
done
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-chunk 16 > gpurun_out/bench_e2e_c16.json 2> gpurun_out/bench_e2e_c16.err; echo "bench rc=$?"
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-chunk 32 > gpurun_out/bench_e2e_c32.json 2> gpurun_out/bench_e2e_c32.err; echo "bench rc=$?"
