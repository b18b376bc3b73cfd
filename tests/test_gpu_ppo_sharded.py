"""PPO mini-batch update partitioned over ranks (BASELINE configs[3]; ACMPC/training_loop.py:134-199):
the all-reduced gradient bucket of two processes, each evaluating half of the mini-batch, must equal the gradient a
single process computes on the whole mini-batch, and so must the parameters after optimiser steps.

Two ranks share cuda:0 here (the GPU test box has one GPU), so the process group is gloo on CUDA tensors — the product
code path (ShardedPPOUpdate, GradBucket on a communication stream) is the one NCCL runs under torchrun
(tools/bench_ppo_update.py)."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu

MB, T, STEPS = 64, 10, 3


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _build(graph, world_rank=None):
    import torch.nn as nn

    from tacdmpc_b200.ACMPC import ActorMPC, ShardedPPOUpdate
    from tacdmpc_b200.dynamics import CartPole
    torch.manual_seed(0)
    f = CartPole()
    actor = ActorMPC(4, 1, T, 0.05, f_dyn=f, f_dyn_jac=f.jac, device="cuda", grad_method="analytic")
    critic = nn.Sequential(nn.Linear(4, 64), nn.Tanh(), nn.Linear(64, 1)).cuda()
    g = torch.Generator(device="cpu").manual_seed(1)
    data = []
    for _ in range(STEPS):
        st = torch.rand(MB, 4, generator=g) * 0.45 - 0.2
        data.append((st, torch.randn(MB, 1, generator=g) * 0.5, torch.randn(MB, generator=g) * 0.1 - 1.0,
                     torch.randn(MB, generator=g), torch.randn(MB, generator=g)))
    first = [d[0][:1].clone() for d in data]          # observation of the global mini-batch's element 0 (quirk Q4)
    if world_rank is not None:
        from tacdmpc_b200.sharding import shard_range
        lo, hi = shard_range(MB, *world_rank)
        data = [tuple(t[lo:hi] for t in d) for d in data]
    data = [d + (f0,) for d, f0 in zip(data, first)]
    data = [tuple(t.cuda() for t in d) for d in data]
    a_opt = torch.optim.Adam(actor.parameters(), lr=1e-3)
    c_opt = torch.optim.Adam(critic.parameters(), lr=1e-3)
    upd = ShardedPPOUpdate(actor, critic, data[0], actor_optimizer=a_opt, critic_optimizer=c_opt, graph=graph)
    return upd, data


def _run(upd, data):
    """gradient bucket after the first step, parameters after all steps"""
    first = None
    for d in data:
        upd(*d)
        if first is None:
            first = upd.bucket.flat.detach().clone().cpu()
    torch.cuda.synchronize()
    return first, torch.cat([p.detach().reshape(-1) for p in upd.params]).cpu()


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(0)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        upd, data = _build(graph=True, world_rank=(rank, world))
        assert upd.world == world and abs(upd.weight - 1.0 / world) < 1e-12
        g, p = _run(upd, data)
        torch.save({"grad": g, "params": p}, os.path.join(out_dir, f"rank{rank}.pt"))
    finally:
        dist.destroy_process_group()


def test_sharded_update_equals_single_process(tmp_path):
    import torch.multiprocessing as mp
    upd, data = _build(graph=False)
    g1, p1 = _run(upd, data)
    assert float(g1.abs().max()) > 1e-4           # the solve's gradient reaches the cost network
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    r0, r1 = (torch.load(tmp_path / f"rank{r}.pt") for r in (0, 1))
    assert torch.equal(r0["grad"], r1["grad"]) and torch.equal(r0["params"], r1["params"])   # ranks stay in lock step
    scale = float(g1.abs().max())
    eg = float((r0["grad"] - g1).abs().max()) / scale
    ep = float((r0["params"] - p1).abs().max())
    print(f"sharded vs single process: gradient max err {eg:.2e} of max |g|, parameters after {STEPS} Adam steps {ep:.2e}")
    # (without ls_state the second rank scores its line searches with ITS first element's cost and ends 7e-5 away: measured)
    # partial sums of a 64-sample mean in a different order: a few ulp of the largest gradient entry.  Adam divides by sqrt(v),
    # which turns a 1e-6 relative gradient difference of a tiny entry into at most one step size (lr = 1e-3) early on: the
    # parameter check is the coarse one (no rank diverged), the gradient check is the parity statement.
    assert eg < 2e-5
    assert ep < 2e-3
