N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29500 bench.py --gpus $N --workload coset --steps 5 --warmup 3 > gpurun_out/bench_coset_n$N.json 2> gpurun_out/bench_coset_n$N.err; echo "coset rc=$?"
tail -3 gpurun_out/bench_coset_n$N.err
cat gpurun_out/bench_coset_n$N.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29501 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_batch_n$N.json 2> gpurun_out/bench_batch_n$N.err; echo "batch rc=$?"
tail -3 gpurun_out/bench_batch_n$N.err
python -c "
import json; d=json.load(open('gpurun_out/bench_batch_n$N.json')); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'])"
