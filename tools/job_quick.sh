for c in 8 16 32; do
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-chunk $c > gpurun_out/bench_e2e_$c.json 2> gpurun_out/bench_e2e_$c.err; echo "bench $c rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_e2e_$c.json')); print($c, d['value'], d['e2e']['value'])"
done
