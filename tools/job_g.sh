TAG=$1
python -m pytest tests/test_convolution.py -m gpu -x -q 2>&1 | tail -2
python tools/conv_time.py > gpurun_out/conv_time_$TAG.json 2>&1; cat gpurun_out/conv_time_$TAG.json
python - <<'PY'
import math, torch, sys, json
sys.path.insert(0, ".")
from algebraist_b200 import _lib
from algebraist_b200.plan import get_plan
dev = torch.device("cuda", 0)
n, B = 10, 128
plan = get_plan(n, torch.float32, dev)
nf = math.factorial(n)
a = torch.randn(B * nf, device=dev); b = torch.randn(B * nf, device=dev); c = torch.empty_like(a)
lib = _lib.load(); st = torch.cuda.current_stream(dev).cuda_stream
for _ in range(2): lib.snfft_block_matmul(plan.handle, a.data_ptr(), b.data_ptr(), c.data_ptr(), B, st)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): lib.snfft_block_matmul(plan.handle, a.data_ptr(), b.data_ptr(), c.data_ptr(), B, st)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
flop = 2 * B * sum(d ** 3 for d in plan.dims)
print(json.dumps({"snfft_block_matmul_ms": ms, "tflops": flop / ms / 1e9}))
PY
