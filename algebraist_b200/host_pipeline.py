"""Host-resident batches: stream them through the GPU with the copies hidden behind the kernels.

The transforms of this package run on device tensors.  When the functions live in (pinned) host memory the
host<->device copies are as long as the transform itself (PCIe: 1.86 GB each way for 128 functions on S_10
against 63 ms of kernels), so `stream_batches` cuts the batch into chunks and runs three stages on three CUDA
streams: chunk i+1 is copied in while chunk i is transformed and chunk i-1 is copied out.

    out = stream_batches(lambda x: sn_ifft(sn_fft(x, n), n), f_host, out_host, chunk=32)

`fn` is any function of a device tensor (rows = functions) that returns a device tensor with the same number of
rows, e.g. a forward+inverse pair, a convolution through the transform, or a projection onto some partitions.
"""
import os

import torch


def bind_host_to_device_node(device=None):
    """Restrict this process to the CPUs of the NUMA node the GPU hangs off, so that host buffers allocated
    (first touched) afterwards are local to the GPU's PCIe root: copies from the far socket run at about half the
    bandwidth.  Call it before allocating the pinned buffers.  Returns the node, or None when the topology cannot
    be read (then nothing is changed)."""
    try:
        idx = torch.cuda.current_device() if device is None else torch.device(device).index or 0
        pr = torch.cuda.get_device_properties(idx)
        bdf = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bdf}/numa_node") as fh:
            node = int(fh.read().strip())
        if node < 0:
            return None
        with open(f"/sys/devices/system/node/node{node}/cpulist") as fh:
            cpus = set()
            for part in fh.read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = cpus & os.sched_getaffinity(0)
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return node
    except (OSError, ValueError, AttributeError, RuntimeError):
        return None


def chunk_schedule(batch: int, chunk: int, ramp: bool = True):
    """Row counts of the chunks.  With `ramp` the first and the last chunks are short (chunk/4, chunk/2): the
    pipeline cannot hide the copy-in of the first chunk nor the copy-out of the last one, so those are kept small
    while the chunks in between stay large enough for the kernels to run at full efficiency."""
    sizes = []
    head = [c for c in (chunk // 4, chunk // 2) if c >= 1] if ramp else []
    if batch < 2 * sum(head) + chunk:
        head = []
    tail = head[::-1]
    left = batch - sum(head) - sum(tail)
    sizes += head
    while left > 0:
        sizes.append(min(chunk, left))
        left -= sizes[-1]
    return sizes + tail


def stream_batches(fn, src_host: torch.Tensor, dst_host: torch.Tensor = None, chunk: int = 32, device=None,
                   ramp: bool = True, nslots: int = 3):
    """Apply `fn` to the rows of `src_host` chunk by chunk on `device`, writing the results to `dst_host`.

    src_host : (B, ...) host tensor; pin it (`.pin_memory()`) or the copies cannot overlap anything.
    dst_host : (B, ...) host tensor for the results (pinned); allocated (pinned) when None, from the shape
               and dtype of the first chunk's result.
    Returns dst_host.  The call returns after the last result has landed in host memory."""
    if not torch.cuda.is_available():
        raise RuntimeError("algebraist_b200 needs a CUDA device (there is no CPU fallback)")
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    batch = src_host.shape[0]
    if chunk < 1:
        raise ValueError("chunk must be positive")
    if batch == 0:
        return dst_host if dst_host is not None else src_host.new_empty(src_host.shape)
    compute = torch.cuda.current_stream(dev)
    s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    nslots = max(2, int(nslots))                       # device-side input / output chunk buffers in flight
    ins = [None] * nslots
    outs = [None] * nslots
    in_ready = [None] * nslots                         # copy-in of the chunk in slot s finished
    in_free = [None] * nslots                          # the transform that read slot s finished
    out_ready = [None] * nslots                        # the transform that produced slot s finished
    out_free = [None] * nslots                         # copy-out of slot s finished
    start = torch.cuda.Event()
    sizes = chunk_schedule(batch, chunk, ramp)
    bounds = [0]
    for m in sizes:
        bounds.append(bounds[-1] + m)
    nchunks = len(sizes)
    cmax = max(sizes)
    # The chunk buffers come from the CURRENT stream's pool of the caching allocator (s_in only starts after
    # `start`, and the call returns after every copy has finished), so repeated calls reuse the same blocks.
    # Allocated under s_in they would belong to a stream that changes from call to call: every call would
    # cudaMalloc its buffers again (profiles r02c: single steps of 100-250 ms).
    for s in range(min(nslots, nchunks)):
        ins[s] = torch.empty((cmax,) + tuple(src_host.shape[1:]), dtype=src_host.dtype, device=dev)
    start.record(compute)
    s_in.wait_event(start)                             # src_host may have been produced by work queued on `compute`

    def copy_in(c):
        s = c % nslots
        lo, hi = bounds[c], bounds[c + 1]
        with torch.cuda.stream(s_in):
            if in_free[s] is not None:
                s_in.wait_event(in_free[s])
            ins[s][:hi - lo].copy_(src_host[lo:hi], non_blocking=True)
            in_ready[s] = torch.cuda.Event()
            in_ready[s].record(s_in)

    for c in range(min(nslots - 1, nchunks)):
        copy_in(c)
    for c in range(nchunks):
        s = c % nslots
        lo, hi = bounds[c], bounds[c + 1]
        if c + nslots - 1 < nchunks:
            copy_in(c + nslots - 1)                    # nslots - 1 chunks ahead: overlaps the transforms before it
        compute.wait_event(in_ready[s])
        if out_free[s] is not None:
            compute.wait_event(out_free[s])            # the previous result in this slot has left the device
        res = fn(ins[s][:hi - lo])
        # The result stays referenced until this slot is reused, and the compute stream waits for the slot's
        # copy-out (out_free) before that, so the block is never handed out again while s_out still reads it:
        # no record_stream (its deferred frees made the caching allocator fall back to cudaMalloc at random
        # points of the pipeline: single steps of 100-250 ms, profiles r02c).
        outs[s] = res
        in_free[s] = torch.cuda.Event()
        in_free[s].record(compute)
        out_ready[s] = torch.cuda.Event()
        out_ready[s].record(compute)
        if dst_host is None:
            dst_host = torch.empty((batch,) + tuple(res.shape[1:]), dtype=res.dtype).pin_memory()
        with torch.cuda.stream(s_out):
            s_out.wait_event(out_ready[s])
            dst_host[lo:hi].copy_(res, non_blocking=True)
            out_free[s] = torch.cuda.Event()
            out_free[s].record(s_out)
    s_out.synchronize()
    return dst_host
