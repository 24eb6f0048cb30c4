"""Kernel timeline of ONE eager step of the bench workload on this rank (torch.profiler / CUPTI -- works under torchrun, where
ncu is not allowed): every kernel of the step in launch order with its device time, averaged over the profiled steps.

    python tools/step_trace.py                                      # 1 GPU
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/step_trace.py   # rank 0 prints
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import bench  # noqa: E402

rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
local_rank = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local_rank)
dev = torch.device("cuda", local_rank)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
workload = sys.argv[1] if len(sys.argv) > 1 else "c3"
cfg = bench.CONFIGS[workload]
B = bench.per_gpu_batch(cfg, "weak", world)
hp = bench.HotPath(cfg, bench.make_inputs(cfg, B, dev, bench.SEED + rank), dev, rank, world)
for _ in range(5):
    hp.step()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
STEPS = 4
from torch.profiler import ProfilerActivity, profile  # noqa: E402

with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(STEPS):
        hp.step()
    torch.cuda.synchronize()
if rank == 0:
    evs = sorted((e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA), key=lambda e: e.time_range.start)
    per = len(evs) // STEPS
    last = evs[-per:]
    t0 = last[0].time_range.start
    rows = []
    for e in last:
        rows.append(dict(start_us=round(e.time_range.start - t0, 1), us=round(e.time_range.end - e.time_range.start, 1), name=e.name[:90]))
    span = last[-1].time_range.end - t0
    print(json.dumps(dict(world=world, workload=workload, kernels_per_step=per, step_span_us=round(span, 1), kernels=rows), indent=0))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
