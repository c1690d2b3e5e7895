#!/usr/bin/env python
"""Benchmark of the S_n FFT hot path (BASELINE.json: S_n FFT+IFFT transforms/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference]
                    [--n 10] [--batch 128] [--dtype f32|f64]

A step is one forward + inverse transform of a batch of synthetic random
functions (torch.randn, like the reference's tests/test_fourier.py:37-41).
Default workload: S_10, batch 128 per GPU, fp32 (BASELINE.json configs[3]).
With N > 1 (torchrun, one rank per GPU) the batch is sharded with no
collective on the data path: every rank transforms its own 128 functions
(weak scaling) and the value is the whole-job aggregate.

Prints ONE JSON line (rank 0).  `--impl reference` times the CPU oracle port of
the reference algorithm (oracle/) on the host cores instead.
"""
import argparse
import ctypes
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "sn_fft_ifft_pairs_per_s"
UNIT = "transforms/s"

# Libraries (NCCL prints its version banner) write to file descriptor 1; the contract is ONE JSON line on
# stdout, so fd 1 is pointed at stderr for the whole run and the line goes to a private copy of stdout.
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(line):
    with os.fdopen(os.dup(_REAL_STDOUT), "w") as fh:
        fh.write(json.dumps(line) + "\n")
        fh.flush()


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--n", type=int, default=10)
    ap.add_argument("--batch", type=int, default=128, help="functions per GPU")
    ap.add_argument("--dtype", default="f32", choices=["f32", "f64"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-numa-bind", action="store_true", help="do not bind the process to the GPU's NUMA node for the e2e leg")
    ap.add_argument("--e2e-slots", type=int, default=3, help="chunk buffers in flight in the end-to-end pipeline")
    ap.add_argument("--e2e-chunk", type=int, default=0, help="functions per chunk of the end-to-end pipeline (0 = batch / 8)")
    ap.add_argument("--workload", default="batch", choices=["batch", "coset"],
                    help="batch: BASELINE configs 2-4 (default); coset: config 5, ONE S_n function (default n = 11) sharded "
                         "over the top-level cosets with a reduce-scatter by partition (torchrun, N >= 2)")
    return ap.parse_args()


def workload_name(a):
    return f"S{a.n} {a.dtype} FFT+IFFT, batch {a.batch} per GPU"


def measured_traffic(kernel, batch, dtype):
    """DRAM bytes per launch of `kernel` from the committed ncu capture (profiles/*traffic*.json, written by
    tools/ncu_traffic.py from `ncu --set full`), or None when no capture matches this workload."""
    best = None
    pdir = os.path.join(ROOT, "profiles")
    try:
        for fn in sorted(os.listdir(pdir)):
            if "traffic" in fn and fn.endswith(".json"):
                with open(os.path.join(pdir, fn)) as fh:
                    for rec in json.load(fh):
                        if rec.get("kernel") == kernel and rec.get("batch") == batch and rec.get("dtype") == dtype:
                            best = {"bytes": rec["dram_bytes_per_launch"], "source": "profiles/" + fn}
    except Exception:
        return None
    return best


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons of one GPU while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], threading.Event()

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout
                parts = [p.strip() for p in out.strip().split(",")]
                if len(parts) >= 7:
                    self.samples.append(parts)
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        self.stop_flag.set()
        self.join(timeout=6)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [nm for i, nm in enumerate(names) if any(s[3 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(self.samples[0][1]),
                "power_w_max": max(float(s[2]) for s in self.samples), "samples": len(self.samples),
                "reasons": reasons}


# --------------------------------------------------------------------------
# CPU side: the oracle port of the reference algorithm
# --------------------------------------------------------------------------
def _oracle_pair(args):
    n, dtype, seed = args
    import numpy as np
    from oracle import sn_oracle as O
    f = np.random.default_rng(seed).standard_normal((1, math.factorial(n))).astype(dtype)
    ft = O.sn_fft_packed_c(f, n)           # oracle level steps with the C inner loops
    inv = O.sn_ifft_packed_c(ft, n)
    return float(abs(inv - f).max())


def cpu_pairs_per_s(n, dtype, nfunc, workers):
    """Forward+inverse of `nfunc` functions on `workers` host processes (oracle port)."""
    import numpy as np
    dt = np.float32 if dtype == "f32" else np.float64
    jobs = [(n, dt, 1234 + i) for i in range(nfunc)]
    t0 = time.perf_counter()
    if workers <= 1:
        for j in jobs:
            _oracle_pair(j)
    else:
        import multiprocessing as mp
        with mp.get_context("fork").Pool(workers) as pool:
            pool.map(_oracle_pair, jobs, chunksize=1)
    return nfunc / (time.perf_counter() - t0)


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    workers = max(1, min(cores, 32))
    from oracle import sn_oracle as O
    O.c_library()
    O.gather_index(a.n)                                  # tables warm (plan time, not step time)
    for k in range(2, a.n + 1):
        for lam in O.generate_partitions(k):
            O._c_tables(lam)
    sample = workers                                      # one function per worker per step
    times = []
    for s in range(a.warmup + a.steps):
        t0 = time.perf_counter()
        cpu_pairs_per_s(a.n, a.dtype, sample, workers)
        if s >= a.warmup:
            times.append(time.perf_counter() - t0)
    total = sum(times)
    value = sample * a.steps / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1000 * total / a.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": a.dtype, "data": "synthetic",
        "config": {"workload": workload_name(a), "n": a.n, "batch_per_gpu": a.batch,
                   "note": "CPU oracle port of the reference algorithm (sparse YOR sweeps, C inner loops); the unmodified "
                           "reference cannot run n >= 9 (SURVEY 0.3) and its recursive inverse raises for n > 5"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": workers, "kind": "port",
                         "sample": f"{sample} functions per step ({workers} processes x 1), forward+inverse"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


# --------------------------------------------------------------------------
# GPU side
# --------------------------------------------------------------------------
def run_native(a):
    import torch
    import torch.distributed as dist
    from algebraist_b200 import _lib, sn_fft, sn_ifft
    from algebraist_b200.plan import get_plan

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    tdt = torch.float32 if a.dtype == "f32" else torch.float64
    esize = 4 if a.dtype == "f32" else 8
    n, B = a.n, a.batch
    nf = math.factorial(n)
    plan = get_plan(n, tdt, dev)
    torch.manual_seed(1000 + rank)
    f = torch.randn(B, nf, dtype=tdt, device=dev)
    F = torch.empty(B * nf, dtype=tdt, device=dev)
    f2 = torch.empty_like(f)
    ws = plan.workspace(B)
    stream = torch.cuda.current_stream(dev)

    def step():
        _lib.check(lib.snfft_forward(plan.handle, f.data_ptr(), B, F.data_ptr(), ws.data_ptr(), ws.numel(),
                                     stream.cuda_stream), "snfft_forward")
        _lib.check(lib.snfft_inverse(plan.handle, F.data_ptr(), B, f2.data_ptr(), ws.data_ptr(), ws.numel(),
                                     stream.cuda_stream), "snfft_inverse")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(a.warmup, 3)):
        step()
    barrier()
    err = ((f2 - f).double().norm() / f.double().norm()).item()

    sampler = ClockSampler(local)
    sampler.start()
    launches0 = lib.snfft_launch_count()
    _lib.check(lib.snfft_profile_begin(plan.handle), "profile_begin")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for _ in range(a.steps):
        step()
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    recs = (_lib.ProfileRecord * 256)()
    nrec = ctypes.c_int()
    _lib.check(lib.snfft_profile_end(plan.handle, recs, 256, ctypes.byref(nrec)), "profile_end")
    launches = lib.snfft_launch_count() - launches0
    clocks = sampler.summary()
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = t.item()
    value = world * B * a.steps / (ms / 1000.0)

    # end to end through the public Python API with host buffers (allocated on the GPU's NUMA node)
    from algebraist_b200 import stream_batches
    from algebraist_b200.host_pipeline import bind_host_to_device_node
    affinity0 = os.sched_getaffinity(0)
    numa_node = None if a.no_numa_bind else bind_host_to_device_node(dev)
    f_host = torch.randn(B, nf, dtype=tdt).pin_memory()
    out_host = torch.empty(B, nf, dtype=tdt).pin_memory()
    e2e_chunk = a.e2e_chunk if a.e2e_chunk > 0 else max(4, B // 8)

    def e2e_step():
        # public API for host-resident batches: chunks of the batch stream through the GPU, the copies of
        # chunk i+1 / i-1 overlap the transform of chunk i (all of them are inside the timed region)
        stream_batches(lambda x: sn_ifft(sn_fft(x, n), n), f_host, out_host, chunk=e2e_chunk, device=dev,
                       nslots=a.e2e_slots)

    for _ in range(3):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(2, min(a.steps, 20))
    e2e_marks = [t0]
    for _ in range(e2e_steps):
        e2e_step()                      # returns when the last result is in host memory
        e2e_marks.append(time.perf_counter())
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = world * B * e2e_steps / e2e_s.item()
    e2e_err = ((out_host - f_host).double().norm() / f_host.double().norm()).item()
    os.sched_setaffinity(0, affinity0)

    if rank == 0:
        peak, peak_src = peaks()
        kinds = {0: "fwd_level", 1: "inv_level", 2: "gather", 3: "scatter", 4: "fwd_base", 5: "inv_base",
                 6: "pack", 7: "unpack"}
        kernels = []
        for i in range(nrec.value):
            r = recs[i]
            ninst = B * (nf // math.factorial(r.level)) if r.kind in (0, 1, 4, 5, 6, 7) else B
            bytes_per_launch = 2 * r.elems_per_instance * ninst * esize
            avg_ms = r.total_ms / r.launches
            kernels.append({"kernel": f"{kinds[r.kind]}_k{r.level}_c{r.kernel_class}", "launches": r.launches,
                            "avg_ms": avg_ms, "share": r.total_ms / ms, "alg_bytes": bytes_per_launch,
                            "alg_gbs": bytes_per_launch / (avg_ms * 1e-3) / 1e9})
        kernels.sort(key=lambda k: -k["share"])
        top = kernels[0]
        alg_bytes_pair = 4 * nf * esize
        traffic = measured_traffic(top["kernel"], B, a.dtype)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps,
            "warmup": max(a.warmup, 3), "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": a.dtype, "data": "synthetic",
            "config": {"workload": workload_name(a), "n": n, "batch_per_gpu": B, "parallelism": f"batch-shard x{world}",
                       "l2_policy": f"inputs larger than L2 ({B * nf * esize / 1e6:.0f} MB per array per GPU)",
                       "round_trip_rel_err": err},
            "roofline": {"bound": "hbm", "kernel": top["kernel"], "achieved": top["alg_gbs"], "peak": peak,
                         "unit": "GB/s", "frac": top["alg_gbs"] / peak,
                         "traffic": traffic["bytes"] if traffic else None,
                         "traffic_source": traffic["source"] if traffic else None,
                         "algorithmic_bytes_per_launch": top["alg_bytes"],
                         "peak_source": peak_src, "kernel_share_of_step": top["share"],
                         "end_to_end_alg_gbs": value / world * alg_bytes_pair / 1e9,
                         "end_to_end_frac": value / world * alg_bytes_pair / 1e9 / peak},
            "kernels": kernels[:40],
            "gpu_launches": int(launches),
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": B * nf * esize,
                    "d2h_bytes_per_step": B * nf * esize, "steps": e2e_steps, "round_trip_rel_err": e2e_err,
                    "ms_per_step_each": [round((y - x) * 1e3, 2) for x, y in zip(e2e_marks, e2e_marks[1:])],
                    "host_numa_node": numa_node,
                    "path": f"pinned host -> stream_batches(sn_ifft(sn_fft(.)), chunk={e2e_chunk}, short first/last chunks) -> pinned host "
                            "(H2D, kernels and D2H of consecutive chunks overlap on three streams)"},
        }
        if world == 1 and not a.no_cpu_baseline:
            nfunc = 4 if n >= 10 else 16
            from oracle import sn_oracle as O
            O.c_library()
            for kk in range(2, n + 1):
                for lam in O.generate_partitions(kk):
                    O._c_tables(lam)
            O.gather_index(n)
            v = cpu_pairs_per_s(n, a.dtype, nfunc, 1)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": 1, "kind": "port",
                                    "sample": f"{nfunc} functions of the same workload, forward+inverse, "
                                              "oracle port (sparse YOR sweeps, C inner loops, 1 thread)"}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def run_coset(a):
    """BASELINE config 5: forward transform of ONE function on S_n, sharded over the top-level cosets;
    the ranks finish with an NCCL reduce-scatter by partition.  Strong scaling: the work is fixed."""
    import torch
    import torch.distributed as dist
    from algebraist_b200 import sn_fft
    from algebraist_b200.distributed import coset_shards, partition_owners, sn_fft_coset_sharded

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev) if world > 1 else dist.init_process_group(
        "nccl", init_method="tcp://127.0.0.1:29531", rank=0, world_size=1, device_id=dev)
    tdt = torch.float32 if a.dtype == "f32" else torch.float64
    esize = 4 if a.dtype == "f32" else 8
    n = a.n if a.n != 10 else 11
    nf = math.factorial(n)
    torch.manual_seed(5)                                  # every rank holds the same function
    f = torch.randn(nf, dtype=tdt, device=dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)     # > L2, written between steps

    def barrier():
        dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(a.warmup, 3)):
        mine = sn_fft_coset_sharded(f, n)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    from algebraist_b200.distributed import sn_fft_coset_partial
    lo_c, hi_c = coset_shards(n, world)[rank]
    times, compute = [], []
    for _ in range(a.steps):
        flush.zero_()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        mine = sn_fft_coset_sharded(f, n)
        e1.record()
        barrier()
        times.append(e0.elapsed_time(e1))
        # the per-rank transform alone (no exchange), for the split between compute and reduce-scatter
        e0.record()
        sn_fft_coset_partial(f, n, range(lo_c, hi_c))
        e1.record()
        barrier()
        compute.append(e0.elapsed_time(e1))
    tc = torch.tensor([sum(compute) / len(compute)], dtype=torch.float64, device=dev)
    dist.all_reduce(tc, op=dist.ReduceOp.MAX)
    t = torch.tensor([sum(times)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = t.item()
    clocks = sampler.summary()
    # parity: the owned partitions against the single-GPU transform of the same function (rank 0 only: 3 x 160 MB)
    err = 0.0
    if rank == 0:
        full = sn_fft(f, n)
        for lam, v in mine.items():
            err = max(err, ((v - full[lam]).double().norm() / full[lam].double().norm().clamp_min(1e-30)).item())
    power = torch.zeros(1, dtype=torch.float64, device=dev)
    for lam, v in mine.items():
        power += v.double().pow(2).sum() * (1 if v.dim() == 0 else v.shape[-1]) / nf
    dist.all_reduce(power)
    plancherel = abs(power.item() / f.double().pow(2).sum().item() - 1.0)
    if rank == 0:
        from algebraist_b200.plan import get_plan
        plan = get_plan(n, tdt, dev)
        owner, bin_size = partition_owners(plan.dims, world)
        line = {
            "metric": "sn_fft_forward_per_s", "value": a.steps / (ms / 1000.0), "unit": UNIT, "n_gpus": world,
            "steps": a.steps, "warmup": max(a.warmup, 3), "ms_per_step": ms / a.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": a.dtype, "data": "synthetic",
            "config": {"workload": f"S{n} {a.dtype} forward FFT of one function, sharded over top-level cosets",
                       "n": n, "parallelism": f"coset-shard x{world} + reduce-scatter by partition",
                       "cosets_per_rank": [hi - lo for lo, hi in coset_shards(n, world)],
                       "reduce_scatter_bytes_per_rank": bin_size * esize * (world - 1),
                       "l2_policy": "256 MB flush buffer written between timed iterations",
                       "compute_ms_max_over_ranks": tc.item(), "rel_err_vs_single_gpu": err,
                       "plancherel_rel_err": plancherel},
            "clocks": clocks,
        }
        emit(line)
    dist.destroy_process_group()


if __name__ == "__main__":
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "coset":
        run_coset(args)
    else:
        run_native(args)
