"""Would a tensor-core path hold the 1e-5 contract of `snfft_block_matmul` (the d x d products of sn_convolve)?

Measured with library GEMMs standing in for a tcgen05 kernel (torch.bmm = cuBLAS; a study tool, not a product path):
  fp32      : torch.bmm with TF32 off (FFMA)                        -- accuracy reference for the contract
  tf32      : torch.bmm with TF32 on (one tensor-core pass, 10-bit mantissa)
  3xtf32    : a = a_hi + a_lo (a_hi = a rounded to TF32), C = a_hi b_hi + a_hi b_lo + a_lo b_hi with TF32 GEMMs
  ours      : snfft_block_matmul (hand-written FFMA tile kernel)
Errors are ||C - C64||_F / ||C64||_F against an fp64 product; times are CUDA events over 20 repetitions.

    python tools/tf32_split_study.py > gpurun_out/tf32_split_study.json
"""
import json
import math
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402


def tf32_round(x):
    """Round-to-nearest-even to 10 mantissa bits (what the tensor core sees)."""
    i = x.view(torch.int32)
    r = ((i >> 13) & 1) + 0x0FFF
    return ((i + r) & ~0x1FFF).view(torch.float32)


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    out = {"batch": 128, "cases": []}
    B = 128
    for d in (35, 128, 288, 768):
        torch.manual_seed(d)
        # Fourier coefficients of random functions on S_10: entries ~ N(0, n!) -- only the scale-free error matters
        a = torch.randn(B, d, d, device="cuda")
        b = torch.randn(B, d, d, device="cuda")
        ref = torch.bmm(a.double(), b.double())
        rec = {"d": d, "flop": 2.0 * B * d ** 3}

        def err(c):
            return ((c.double() - ref).norm() / ref.norm()).item()

        torch.backends.cuda.matmul.allow_tf32 = False
        rec["fp32_err"] = err(torch.bmm(a, b))
        rec["fp32_ms"] = timed(lambda: torch.bmm(a, b))
        torch.backends.cuda.matmul.allow_tf32 = True
        rec["tf32_err"] = err(torch.bmm(a, b))
        rec["tf32_ms"] = timed(lambda: torch.bmm(a, b))
        ah, bh = tf32_round(a), tf32_round(b)
        al, bl = a - ah, b - bh

        def split():
            return torch.bmm(ah, bh) + (torch.bmm(ah, bl) + torch.bmm(al, bh))

        rec["3xtf32_err"] = err(split())
        rec["3xtf32_ms"] = timed(split)
        torch.backends.cuda.matmul.allow_tf32 = False
        for k in ("fp32", "tf32", "3xtf32"):
            rec[k + "_tflops"] = rec["flop"] / rec[k + "_ms"] / 1e9
        out["cases"].append(rec)
    # the product's kernel on the real shapes: all 42 partitions of S_10, batch 128, one launch
    import algebraist_b200 as A
    from algebraist_b200 import _lib
    from algebraist_b200.plan import get_plan
    n = 10
    plan = get_plan(n, torch.float32, torch.device("cuda", 0))
    pa = torch.randn(B * math.factorial(n), device="cuda")
    pb = torch.randn(B * math.factorial(n), device="cuda")
    pc = torch.empty_like(pa)
    lib = _lib.load()
    st = torch.cuda.current_stream().cuda_stream
    ms = timed(lambda: _lib.check(lib.snfft_block_matmul(plan.handle, pa.data_ptr(), pb.data_ptr(), pc.data_ptr(), B, st), "mm"), reps=5)
    flop = 2.0 * B * sum(d ** 3 for d in plan.dims)
    worst = 0.0
    for d, off in zip(plan.dims, plan.offsets):
        if d in (1, 35, 288, 768):
            x = pa[off * B:(off + d * d) * B].view(B, d, d)
            y = pb[off * B:(off + d * d) * B].view(B, d, d)
            z = pc[off * B:(off + d * d) * B].view(B, d, d)
            r = torch.bmm(x.double(), y.double())
            worst = max(worst, ((z.double() - r).norm() / r.norm()).item())
    out["snfft_block_matmul"] = {"n": n, "ms": ms, "tflops": flop / ms / 1e9, "worst_rel_err": worst, "launches": 1}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
