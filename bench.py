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
    ap.add_argument("--no-sub-records", action="store_true",
                    help="skip the S9 x 1024 / S10 fp64 sub-records and the S11 coset record of the default workload")
    ap.add_argument("--no-numa-bind", action="store_true", help="do not bind the process to the GPU's NUMA node for the e2e leg")
    ap.add_argument("--e2e-slots", type=int, default=3, help="chunk buffers in flight in the end-to-end pipeline")
    ap.add_argument("--e2e-chunk", type=int, default=0, help="functions per chunk of the end-to-end pipeline (0 = batch / 16)")
    ap.add_argument("--workload", default="batch", choices=["batch", "coset"],
                    help="batch: BASELINE configs 2-4 (default); coset: config 5, ONE S_n function (default n = 11) sharded "
                         "over the top-level cosets with a reduce-scatter by partition (torchrun, N >= 2)")
    return ap.parse_args()


def workload_name(a):
    return f"S{a.n} {a.dtype} FFT+IFFT, batch {a.batch} per GPU"


def measured_traffic(kernel, batch, dtype, n=10):
    """DRAM bytes per launch of `kernel` from the committed ncu capture of THIS round's kernels
    (profiles/r02* / r03* ... traffic*.json, written by tools/ncu_traffic.py from `ncu --set full`; each record names the
    capture and the commit it was taken on), or None when no capture of the current kernels matches."""
    best = None
    pdir = os.path.join(ROOT, "profiles")
    try:
        for fn in sorted(os.listdir(pdir)):
            if fn[:2] == "r0" and fn[2:3] in "23456789" and "traffic" in fn and fn.endswith(".json"):
                with open(os.path.join(pdir, fn)) as fh:
                    for rec in json.load(fh):
                        if (rec.get("kernel") == kernel and rec.get("batch") == batch and rec.get("dtype") == dtype
                                and rec.get("n", 10) == n):
                            best = {"bytes": rec["dram_bytes_per_launch"],
                                    "source": f"profiles/{fn} (capture {rec.get('capture')}, commit {rec.get('commit')})"}
    except Exception:
        return None
    return best


def cpu_model():
    try:
        with open("/proc/cpuinfo") as fh:
            for ln in fh:
                if ln.startswith("model name"):
                    return ln.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


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
        "config": config_of(a, a.gpus),
        "note": "CPU oracle port of the reference algorithm (sparse YOR sweeps, C inner loops); the unmodified "
                "reference cannot run n >= 9 (SURVEY 0.3) and its recursive inverse raises for n > 5",
        "host": {"cpu_model": cpu_model(), "cpus": os.cpu_count()},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": workers, "kind": "port",
                         "sample": f"{sample} functions per step ({workers} processes x 1), forward+inverse"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


# --------------------------------------------------------------------------
# GPU side
# --------------------------------------------------------------------------
KINDS = {0: "fwd_level", 1: "inv_level", 2: "gather", 3: "scatter", 4: "fwd_base", 5: "inv_base", 6: "pack", 7: "unpack"}


def config_of(a, world):
    """The workload description; the native and the reference arm emit the same keys and values."""
    nf = math.factorial(a.n)
    esize = 4 if a.dtype == "f32" else 8
    return {"workload": workload_name(a), "n": a.n, "batch_per_gpu": a.batch, "parallelism": f"batch-shard x{world}",
            "l2_policy": f"inputs larger than L2 ({a.batch * nf * esize / 1e6:.0f} MB per array per GPU)"}


def device_record(lib, n, B, dtype, steps, warmup, world, dev, barrier, all_reduce_max):
    """Device-timed forward + inverse of B resident functions per rank: whole-job value, per-launch CUDA-event
    timings of every kernel and the roofline of the dominant one."""
    import torch
    from algebraist_b200 import _lib
    from algebraist_b200.plan import get_plan
    tdt = torch.float32 if dtype == "f32" else torch.float64
    esize = 4 if dtype == "f32" else 8
    nf = math.factorial(n)
    plan = get_plan(n, tdt, dev)
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

    for _ in range(max(warmup, 3)):
        step()
    barrier()
    err = ((f2 - f).double().norm() / f.double().norm()).item()
    launches0 = lib.snfft_launch_count()
    _lib.check(lib.snfft_profile_begin(plan.handle), "profile_begin")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for _ in range(steps):
        step()
    e1.record(stream)
    barrier()
    ms = all_reduce_max(e0.elapsed_time(e1))
    recs = (_lib.ProfileRecord * 256)()
    nrec = ctypes.c_int()
    _lib.check(lib.snfft_profile_end(plan.handle, recs, 256, ctypes.byref(nrec)), "profile_end")
    launches = lib.snfft_launch_count() - launches0
    value = world * B * steps / (ms / 1000.0)
    peak, peak_src = peaks()
    kernels = []
    for i in range(nrec.value):
        r = recs[i]
        ninst = B * (nf // math.factorial(r.level)) if r.kind in (0, 1, 4, 5, 6, 7) else B
        bytes_per_launch = 2 * r.elems_per_instance * ninst * esize
        avg_ms = r.total_ms / r.launches
        kernels.append({"kernel": f"{KINDS[r.kind]}_k{r.level}_c{r.kernel_class}", "launches": r.launches,
                        "avg_ms": avg_ms, "share": r.total_ms / ms, "alg_bytes": bytes_per_launch,
                        "alg_gbs": bytes_per_launch / (avg_ms * 1e-3) / 1e9})
    kernels.sort(key=lambda k: -k["share"])
    top = kernels[0]
    alg_bytes_pair = 4 * nf * esize
    traffic = measured_traffic(top["kernel"], B, dtype, n)
    del f, F, f2, ws
    return {
        "value": value, "ms_per_step": ms / steps, "round_trip_rel_err": err, "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "kernel": top["kernel"], "achieved": top["alg_gbs"], "peak": peak,
                     "unit": "GB/s", "frac": top["alg_gbs"] / peak,
                     "traffic": traffic["bytes"] if traffic else None,
                     "traffic_source": traffic["source"] if traffic else None,
                     "algorithmic_bytes_per_launch": top["alg_bytes"],
                     "peak_source": peak_src, "kernel_share_of_step": top["share"],
                     "end_to_end_alg_gbs": value / world * alg_bytes_pair / 1e9,
                     "end_to_end_frac": value / world * alg_bytes_pair / 1e9 / peak,
                     "kernel_classes": "c-3 column kernels, c-4 Z kernel, c-5 T kernel (top / bottom split), c-7 top-level T kernel on the packed "
                                       "layout (pack / unpack fused), c-2 vector "
                                       "gather / pack, c-1 generated levels 2..5, c-8 levels 2..6 in one kernel (fp32), c>=0 sweep kernels"},
        "kernels": kernels[:40],
    }


def coset_record(n, dtype, steps, world, rank, dev, barrier, all_reduce_max):
    """BASELINE config 5 inside the default run: forward transform of ONE S_n function, sharded over the
    n (n-1) two-level cosets with a reduce-scatter by partition (N >= 2), or the plain transform (N = 1)."""
    import torch
    import torch.distributed as dist
    from algebraist_b200 import sn_fft
    from algebraist_b200.distributed import (_reduce_scatter_sum, coset_pair_shards, partition_owners,
                                             sn_fft_coset2_partial, sn_fft_coset_sharded)
    from algebraist_b200.plan import get_plan
    tdt = torch.float32 if dtype == "f32" else torch.float64
    esize = 4 if dtype == "f32" else 8
    nf = math.factorial(n)
    gen = torch.Generator(device=dev)
    gen.manual_seed(5)                                    # every rank holds the same function
    f = torch.randn(nf, dtype=tdt, device=dev, generator=gen)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)     # > L2, written between steps
    run = (lambda: sn_fft_coset_sharded(f, n)) if world > 1 else (lambda: sn_fft(f, n))
    for _ in range(3):
        mine = run()
    barrier()
    times, compute, coll = [], [], []
    plan = get_plan(n, tdt, dev)
    owner, bin_size = partition_owners(plan.dims, world)
    pairs = coset_pair_shards(n, world)[rank]
    for _ in range(steps):
        flush.zero_()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        mine = run()
        e1.record()
        barrier()
        times.append(e0.elapsed_time(e1))
        if world > 1:                                     # the split: per-rank transform alone, collective alone
            e0.record()
            sn_fft_coset2_partial(f, n, pairs)
            e1.record()
            barrier()
            compute.append(e0.elapsed_time(e1))
            send = torch.zeros(world * bin_size, dtype=tdt, device=dev)
            recv = torch.empty(bin_size, dtype=tdt, device=dev)
            barrier()
            e0.record()
            _reduce_scatter_sum(send, recv)
            e1.record()
            barrier()
            coll.append(e0.elapsed_time(e1))
    ms = all_reduce_max(sum(times)) / steps
    rec = {"metric": "sn_fft_forward_per_s", "value": 1000.0 / ms, "unit": UNIT, "n_gpus": world, "ms_per_step": ms,
           "scaling": "strong", "dtype": dtype,
           "config": {"workload": f"S{n} {dtype} forward FFT of one function" + (", sharded over the two-level cosets" if world > 1 else ""),
                      "n": n, "parallelism": f"coset-pair-shard x{world} + reduce-scatter by partition" if world > 1 else "single GPU",
                      "l2_policy": "256 MB flush buffer written between timed iterations"}}
    if world > 1:
        rec["units_per_rank"] = [len(p) for p in coset_pair_shards(n, world)]
        rec["compute_ms_max_over_ranks"] = all_reduce_max(sum(compute)) / steps
        rec["collective"] = {"op": "ncclReduceScatter(sum) over the packed coefficients, binned by partition",
                             "bytes_sent_per_rank": bin_size * esize * (world - 1),
                             "ms_max_over_ranks": all_reduce_max(sum(coll)) / steps}
        # parity: the owned partitions against the single-GPU transform of the same function (rank 0: 3 x 160 MB)
        err = 0.0
        if rank == 0:
            full = sn_fft(f, n)
            for lam, v in mine.items():
                err = max(err, ((v - full[lam]).double().norm() / full[lam].double().norm().clamp_min(1e-30)).item())
        rec["rel_err_vs_single_gpu"] = err
        power = torch.zeros(1, dtype=torch.float64, device=dev)
        for lam, v in mine.items():
            power += v.double().pow(2).sum() * (1 if v.dim() == 0 else v.shape[-1]) / nf
        dist.all_reduce(power)
        rec["plancherel_rel_err"] = abs(power.item() / f.double().pow(2).sum().item() - 1.0)
    del flush
    return rec


def run_native(a):
    import torch
    import torch.distributed as dist
    from algebraist_b200 import _lib, sn_fft, sn_ifft

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
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
    torch.manual_seed(1000 + rank)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def all_reduce_max(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    sampler = ClockSampler(local)
    sampler.start()
    main = device_record(lib, n, B, a.dtype, a.steps, a.warmup, world, dev, barrier, all_reduce_max)
    clocks = sampler.summary()

    # the other BASELINE configurations of the metric, measured the same way in the same process
    subs = []
    default_workload = (n, B, a.dtype) == (10, 128, "f32")
    if default_workload and not a.no_sub_records:
        for sn, sb, sdt in ((9, 1024, "f32"), (10, 128, "f64")):
            r = device_record(lib, sn, sb, sdt, max(3, min(a.steps, 5)), 3, world, dev, barrier, all_reduce_max)
            r.update({"metric": METRIC, "unit": UNIT, "n_gpus": world, "dtype": sdt,
                      "config": {"workload": f"S{sn} {sdt} FFT+IFFT, batch {sb} per GPU", "n": sn, "batch_per_gpu": sb}})
            r["kernels"] = r["kernels"][:8]
            subs.append(r)
        torch.cuda.empty_cache()

    # end to end through the public Python API with host buffers: every rank gets its own CPUs (the GPU's NUMA
    # node when the box exposes one, else an equal slice of the allowed CPUs) before it allocates its pinned buffers
    from algebraist_b200 import stream_batches
    from algebraist_b200.host_pipeline import bind_host_to_device_node
    affinity0 = os.sched_getaffinity(0)
    numa_node = None if a.no_numa_bind else bind_host_to_device_node(dev)
    cpu_slice = None
    if numa_node is None and local_world > 1 and not a.no_numa_bind:
        cpus = sorted(affinity0)
        per = max(1, len(cpus) // local_world)
        cpu_slice = cpus[local * per:(local + 1) * per] or cpus
        os.sched_setaffinity(0, cpu_slice)
    f_host = torch.randn(B, nf, dtype=tdt).pin_memory()
    out_host = torch.empty(B, nf, dtype=tdt).pin_memory()
    e2e_chunk = a.e2e_chunk if a.e2e_chunk > 0 else max(4, B // 16)

    # copy floor of this box with all N ranks copying at once: the step's input host -> device and its output
    # device -> host on two streams, nothing else (what an ideal pipeline would leave)
    d_in = torch.empty(B, nf, dtype=tdt, device=dev)
    d_out = torch.empty(B, nf, dtype=tdt, device=dev)
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    floors = []
    for _ in range(4):
        barrier()
        t0 = time.perf_counter()
        with torch.cuda.stream(s1):
            d_in.copy_(f_host, non_blocking=True)
        with torch.cuda.stream(s2):
            out_host.copy_(d_out, non_blocking=True)
        s1.synchronize()
        s2.synchronize()
        floors.append(all_reduce_max((time.perf_counter() - t0) * 1e3))
    del d_in, d_out
    pcie_floor_ms = min(floors[1:])

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
    e2e_s = all_reduce_max(time.perf_counter() - t0)
    e2e_value = world * B * e2e_steps / e2e_s
    e2e_err = ((out_host - f_host).double().norm() / f_host.double().norm()).item()
    os.sched_setaffinity(0, affinity0)
    del f_host, out_host
    torch.cuda.empty_cache()

    coset = None
    if default_workload and not a.no_sub_records:
        coset = coset_record(11, "f32", 5, world, rank, dev, barrier, all_reduce_max)

    if rank == 0:
        cfg = config_of(a, world)
        line = {
            "metric": METRIC, "value": main["value"], "unit": UNIT, "n_gpus": world, "steps": a.steps,
            "warmup": max(a.warmup, 3), "ms_per_step": main["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": a.dtype, "data": "synthetic",
            "config": cfg,
            "checks": {"round_trip_rel_err": main["round_trip_rel_err"]},
            "roofline": main["roofline"],
            "kernels": main["kernels"],
            "gpu_launches": main["gpu_launches"],
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": B * nf * esize,
                    "d2h_bytes_per_step": B * nf * esize, "steps": e2e_steps, "round_trip_rel_err": e2e_err,
                    "ms_per_step": 1e3 * e2e_s / e2e_steps,
                    "ms_per_step_each": [round((y - x) * 1e3, 2) for x, y in zip(e2e_marks, e2e_marks[1:])],
                    "pcie_floor_ms": pcie_floor_ms,
                    "pcie_floor_note": f"measured in this run: all {world} rank(s) copy the step's input H2D and its output D2H at "
                                       "once from / to pinned memory, max over ranks, best of 3",
                    "vs_pcie_floor": (1e3 * e2e_s / e2e_steps) / pcie_floor_ms,
                    "host_numa_node": numa_node, "cpu_slice": cpu_slice,
                    "path": f"pinned host -> stream_batches(sn_ifft(sn_fft(.)), chunk={e2e_chunk}, short first/last chunks) -> pinned host "
                            "(H2D, kernels and D2H of consecutive chunks overlap on three streams)"},
            "host": {"cpu_model": cpu_model(), "cpus": os.cpu_count()},
        }
        if subs:
            line["sub_records"] = subs
        if coset:
            line["coset"] = coset
        if world == 1 and not a.no_cpu_baseline:
            nfunc = 4 if n >= 10 else 16
            from oracle import sn_oracle as O
            O.c_library()
            for kk in range(2, n + 1):
                for lam in O.generate_partitions(kk):
                    O._c_tables(lam)
            O.gather_index(n)
            v = cpu_pairs_per_s(n, a.dtype, nfunc, 1)
            cores = os.cpu_count() or 1
            workers = max(1, min(cores, 32))
            vall = cpu_pairs_per_s(n, a.dtype, workers, workers)
            line["cpu_baseline"] = {"value": vall, "unit": UNIT, "cores": workers, "kind": "port",
                                    "sample": f"{workers} functions of the same workload on {workers} processes, forward+inverse, "
                                              "oracle port (sparse YOR sweeps, C inner loops)",
                                    "one_thread": {"value": v, "sample": f"{nfunc} functions, 1 thread"},
                                    "cpu_model": cpu_model(),
                                    "reference_unmodified": reference_here()}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def reference_here():
    """Timings of the UNMODIFIED reference taken in the build container by
    tools/time_reference_here.py -- the reference cannot travel to the GPU box -- or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "r02_reference_cpu_build_container.json")) as fh:
            return json.load(fh)
    except Exception:
        return None


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
