"""PPO update phase on N GPUs (VERDICT r1 row E2 / BASELINE configs[3]; ACMPC/training_loop.py:134-199): the mini-batch
partitioned over ranks, per rank one CUDA-graph replay (cost network + iLQR solve + losses + backward), ONE flat NCCL
all-reduce of the actor + critic gradients on a communication stream, two fused Adam steps.

    python tools/bench_ppo_update.py                                            # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
        tools/bench_ppo_update.py

Weak scaling: the reference's mini-batch of 512 (and 4096) samples PER RANK; strong: a global mini-batch of 4096 split.
Times are CUDA events per update step, median over the steps, max over ranks.  The critic is a stand-in MLP with the
parameter count of the reference's CriticTransformer configuration (~0.6 M; the critic model itself is out of scope)."""
import json
import os
import sys

import torch
import torch.distributed as dist
import torch.nn as nn

sys.path.insert(0, ".")
from tacdmpc_b200.ACMPC import ActorMPC, ShardedPPOUpdate  # noqa: E402
from tacdmpc_b200.dynamics import CartPole  # noqa: E402

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
T, STEPS, WARM = 20, 30, 5


def one(mb_local, label):
    torch.manual_seed(0)
    f = CartPole()
    actor = ActorMPC(4, 1, T, 0.05, f_dyn=f, f_dyn_jac=f.jac, device=str(dev), grad_method="analytic")
    critic = nn.Sequential(nn.Linear(4, 768), nn.Tanh(), nn.Linear(768, 768), nn.Tanh(), nn.Linear(768, 1)).to(dev)
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    mk = lambda: (torch.rand(mb_local, 4, device=dev, generator=g) * 0.45 - 0.2, torch.randn(mb_local, 1, device=dev, generator=g),
                  torch.randn(mb_local, device=dev, generator=g) * 0.1 - 1.0, torch.randn(mb_local, device=dev, generator=g),
                  torch.randn(mb_local, device=dev, generator=g), torch.full((1, 4), 0.01, device=dev))
    batches = [mk() for _ in range(4)]
    a_opt = torch.optim.Adam(actor.parameters(), lr=3e-4, fused=True)
    c_opt = torch.optim.Adam(critic.parameters(), lr=1e-3, fused=True)
    upd = ShardedPPOUpdate(actor, critic, batches[0], actor_optimizer=a_opt, critic_optimizer=c_opt)
    for i in range(WARM):
        upd(*batches[i % 4])
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(STEPS)]
    for i, (a, b, c) in enumerate(ev):
        a.record()
        upd(*batches[i % 4])
        c.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(c) for a, _, c in ev)
    ar = sorted(a.elapsed_time(upd.bucket.done) for a, _, _ in ev[-1:])   # last step: start -> all-reduce done
    med = torch.tensor([ts[len(ts) // 2]], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(med, op=dist.ReduceOp.MAX)
    # every rank must hold identical parameters after the run (the all-reduced gradient is the same everywhere)
    chk = torch.cat([p.detach().reshape(-1) for p in upd.params]).double().sum().reshape(1)
    lo, hi = chk.clone(), chk.clone()
    if world > 1:
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    if rank == 0:
        ms = float(med.item())
        print(json.dumps({"what": label, "n_gpus": world, "mini_batch_per_rank": mb_local, "mini_batch_global": mb_local * world,
                          "ms_per_update": round(ms, 4), "samples_per_s": round(mb_local * world / ms * 1e3, 1),
                          "grad_bucket_bytes": upd.bucket.flat.numel() * 4, "start_to_allreduce_done_ms": round(ar[0], 4),
                          "mean_iters": round(float(actor.mpc.iters_last.float().mean()), 2),
                          "params_identical_across_ranks": bool(lo.item() == hi.item())}), flush=True)
    del upd, actor, critic


for mb in (512, 4096):
    one(mb, "weak: mini-batch per rank fixed")
if world > 1:
    one(4096 // world, "strong: global mini-batch 4096")
    dist.destroy_process_group()
