"""2-GPU microbenchmark of the all-gather building blocks (run under torchrun --nproc-per-node 2).

Measures, with and without a concurrent stream of tcgen05 GEMMs on the compute stream, the outbound
rate of (a) nrb_push_rows (SM store kernel) and (b) nrb_push_bytes (copy engines) when every chunk is
sent `fan` times to the peer -- `fan`=7 reproduces the per-GPU NVLink volume of an 8-GPU all-gather.
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem

from news_recommendation_project_v2_b200 import _lib, ops

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
d, rows, n_chunks = 1024, 32768, 16
full = symm_mem.empty((rows * n_chunks, d), dtype=torch.bfloat16, device=dev)
hdl = symm_mem.rendezvous(full, dist.group.WORLD)
peer = hdl.buffer_ptrs[(rank + 1) % world]
peers = [hdl.buffer_ptrs[(rank + 1 + k) % world] for k in range(world - 1)]  # rotated: no common target
src = torch.randn(rows * n_chunks, d, device=dev).to(torch.bfloat16)
a = torch.randn(32768, 4096, device=dev).to(torch.bfloat16)
w = torch.randn(4096, 4096, device=dev).to(torch.bfloat16)
bias = torch.zeros(4096, device=dev)
res = []


def gemms(n):
    for _ in range(n):
        ops.linear(a, w, bias, _lib.EPI_RELU, None, torch.bfloat16)


def dests(fan):
    """world == 2: the one peer `fan` times (emulates the volume); world > 2: the real peers."""
    return [peer] * fan if world == 2 else peers[:fan]


def run(kind, fan, n_streams, with_gemm):
    comms = [torch.cuda.Stream(priority=-1) for _ in range(n_streams)]
    compute = torch.cuda.current_stream()
    torch.cuda.synchronize(); dist.barrier()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for c in range(n_chunks):
        chunk = src[c * rows:(c + 1) * rows]
        ev = torch.cuda.Event(); ev.record(compute)
        for st in comms:
            st.wait_event(ev)
        if kind == "sm":
            with torch.cuda.stream(comms[0]):
                ops.push_rows(chunk, dests(fan), torch.bfloat16, c * rows, d)
        elif kind == "dma":
            for k in range(fan):
                with torch.cuda.stream(comms[k % n_streams]):
                    ops.push_bytes(chunk, [dests(fan)[k]], c * rows * d * 2)
        if with_gemm:
            gemms(2)  # ~1.5 ms of tensor work per chunk
    for st in comms:
        ev = torch.cuda.Event(); ev.record(st); compute.wait_event(ev)
    t1.record()
    torch.cuda.synchronize(); dist.barrier()
    ms = t0.elapsed_time(t1)
    gb = 0 if kind == "none" else rows * d * 2 * fan * n_chunks / 1e9
    r = dict(kind=kind, fan=fan, streams=n_streams, with_gemm=with_gemm, ms=round(ms, 3),
             out_gbs=round(gb / ms * 1e3, 1))
    if rank == 0:
        print(r, flush=True)
        res.append(r)


gemms(3)
for with_gemm in (False, True):
    run("none", 0, 1, with_gemm)
    for fan in ((1, 7) if world == 2 else (1, world - 1)):
        run("sm", fan, 1, with_gemm)
        for ns in (1, 2, 4, 7):
            if ns <= fan:
                run("dma", fan, ns, with_gemm)
if rank == 0:
    json.dump(res, open("gpurun_out/push_bench_n%d.json" % world, "w"), indent=1)
dist.barrier()
dist.destroy_process_group()
