TAG=$1
N=$2
python tools/kernel_times.py --n 11 --batch 1 --forward-only
python -m pytest tests/test_gpu_multirank.py -m gpu -x -q 2>&1 | tail -5
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_${TAG}_n$N.json 2> gpurun_out/bench_${TAG}_n$N.err; echo "bench rc=$?"
tail -3 gpurun_out/bench_${TAG}_n$N.err
python tools/show_bench.py gpurun_out/bench_${TAG}_n$N.json | head -8
