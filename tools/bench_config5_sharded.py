"""BASELINE configs[4] as stated (LTI sweep nx in {8,16,32}, nu in {4,8}, T=50, B = 32 768 over the GPUs of one node): the batch is
partitioned over ranks (no data-path collective), quirk Q4's line-search cost is GLOBAL element 0's, broadcast from
rank 0 once; forward + backward are device-timed per rank, max over ranks (developer tool; the driver's contract line is bench.py).

    python tools/bench_config5_sharded.py                       # 1 GPU, B = 4096 (the per-GPU share of 32 768 / 8)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 \
        tools/bench_config5_sharded.py --total 32768
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, ".")
from tacdmpc_b200 import ops, workloads as WL  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--total", type=int, default=0, help="global batch (default: 4096 per rank)")
ap.add_argument("--reps", type=int, default=10)
args = ap.parse_args()
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
total = args.total or 4096 * world
Bl = total // world
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)

for nx, nu in ((8, 4), (16, 4), (16, 8), (32, 8)):
    w = WL.lti(nx, nu, B=Bl, seed=rank)          # A, B come from the same generator state on every rank only for seed 0:
    w0 = WL.lti(nx, nu, B=1, seed=0)             # take the dynamics (and nothing else) from the seed-0 draw
    w["spec"]["A"], w["spec"]["B"] = w0["spec"]["A"], w0["spec"]["B"]
    s = WL.gpu_spec(w["spec"], dev)
    d = {k: v.to(dev) for k, v in w["inp"].items()}
    ls = [t.to(dev).contiguous() for t in w["ls"]]
    if world > 1:                                # quirk Q4: global element 0 = rank 0's element 0
        flat = torch.cat([t.reshape(-1) for t in ls])
        dist.broadcast(flat, src=0)
        o, out = 0, []
        for t in ls:
            out.append(flat[o:o + t.numel()].view_as(t).contiguous())
            o += t.numel()
        ls = out
    ls = tuple(ls)
    gX = torch.zeros(Bl, s.T + 1, nx, device=dev)
    gU = torch.zeros(Bl, s.T, nu, device=dev)
    gU[:, 0] = 1

    def step():
        fw = ops.ilqr_forward(s, d["x0"], d["C"], d["c"], d["Cf"], d["cf"], d["xref"], d["uref"], d["Uinit"], ls_cost=ls, want_trace=False)
        ops.ilqr_backward(s, fw["X"], fw["U"], d["C"], d["xref"], d["uref"], fw["tight"], gX, gU, Cf_batched=True)
        return fw

    for _ in range(3):
        fw = step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ts = []
    for _ in range(args.reps):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fw = step()
        b.record()
        b.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    med = torch.tensor([ts[len(ts) // 2]], device=dev, dtype=torch.float64)
    its = fw["iters"].double().sum().reshape(1)
    if world > 1:
        dist.all_reduce(med, op=dist.ReduceOp.MAX)
        dist.all_reduce(its)
    if rank == 0:
        ms = float(med.item())
        print(json.dumps({"config": f"5 LTI nx={nx} nu={nu} T=50", "n_gpus": world, "B_total": total, "B_per_gpu": Bl, "ms_per_step_max_over_ranks": round(ms, 4),
                          "solves_per_s": round(total / ms * 1e3, 1), "mean_iters": round(float(its.item()) / total, 3)}), flush=True)
    del d, fw, w
    torch.cuda.empty_cache()
if world > 1:
    dist.destroy_process_group()
