"""Pure pinned host->device copy bandwidth at 1 / 2 / 4 / 8 ranks: is the end-to-end ceiling of
bench.py's float32 host path the box's host memory system or the way gat_transcribe_clips_host chunks its copies?

    python tools/h2d_probe.py                                     # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/h2d_probe.py

Per rank: the bench's payload (4096 x 22050 float32 = 361 MB, pinned) copied to the GPU `reps` times, (a) as ONE
cudaMemcpyAsync, (b) in the 592-clip chunks gat_transcribe_clips_host uses, alternating two device buffers, (c) in
chunks on two streams.  All ranks start together (barrier) and the aggregate is total bytes / slowest rank's time
(CUDA events).  Rank 0 prints one JSON object.
"""
import json, os, sys
import torch

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); lr = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)
N, n, reps = 4096, 22050, 10
host = torch.empty((N, n), dtype=torch.float32).pin_memory()
host.normal_()
whole = torch.empty_like(host, device=dev)
chunk = 592
bufs = [torch.empty((chunk, n), dtype=torch.float32, device=dev) for _ in range(2)]
streams = [torch.cuda.Stream(dev) for _ in range(2)]


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)


def timed(fn):
    fn(); barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    for s in streams:
        torch.cuda.current_stream(dev).wait_stream(s)
    e1.record(); barrier()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = float(t.item())
    return world * reps * host.numel() * 4 / (ms * 1e-3) / 1e9, ms / reps


def one_copy():
    whole.copy_(host, non_blocking=True)


def chunked_one_stream():
    for k, c0 in enumerate(range(0, N, chunk)):
        c1 = min(N, c0 + chunk)
        bufs[k & 1][: c1 - c0].copy_(host[c0:c1], non_blocking=True)


def chunked_two_streams():
    cur = torch.cuda.current_stream(dev)
    for s in streams:
        s.wait_stream(cur)
    for k, c0 in enumerate(range(0, N, chunk)):
        c1 = min(N, c0 + chunk)
        with torch.cuda.stream(streams[k & 1]):
            bufs[k & 1][: c1 - c0].copy_(host[c0:c1], non_blocking=True)


out = {"ranks": world, "bytes_per_rank": host.numel() * 4, "reps": reps}
for name, fn in (("one_copy", one_copy), ("chunks_592_one_stream", chunked_one_stream), ("chunks_592_two_streams", chunked_two_streams)):
    gbs, ms = timed(fn)
    out[name] = {"aggregate_GBps": round(gbs, 1), "per_rank_GBps": round(gbs / world, 1), "ms_per_361MB": round(ms, 3)}
try:
    import pynvml
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(lr)
    out["pcie_rank0"] = {"gen": pynvml.nvmlDeviceGetCurrPcieLinkGeneration(h), "width": pynvml.nvmlDeviceGetCurrPcieLinkWidth(h)}
except Exception as e:
    out["pcie_rank0"] = f"unavailable ({type(e).__name__})"
out["cpus_allowed_rank0"] = len(os.sched_getaffinity(0))
if rank == 0:
    print(json.dumps(out))
if world > 1:
    dist.destroy_process_group()
