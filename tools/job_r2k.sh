TAG=$1
N=$2
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_${TAG}_n$N.json 2> gpurun_out/bench_${TAG}_n$N.err; echo "bench rc=$?"
tail -3 gpurun_out/bench_${TAG}_n$N.err
python tools/show_bench.py gpurun_out/bench_${TAG}_n$N.json 2>/dev/null | head -6
python -m pytest tests/test_gpu_multirank.py -m gpu -x -q 2>&1 | tail -3
nvidia-smi topo -m 2>/dev/null | head -14
lscpu | grep -E "Model name|^CPU\(s\)|NUMA"
