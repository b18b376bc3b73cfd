"""SURVEY §8d config 4, end to end through the ActorMPC call site (developer tool): (a) the rollout-phase policy call
(forward only, no grad) at B = 4096 and (b) the PPO update's evaluate_actions + backward through the cost network at
the reference's mini-batch B = 512 and at B = 4096 — compact-diagonal (fused) vs dense cost formulation, with the
share of the MPC solve kernels in each.

    python tools/bench_actor.py
"""
import sys

import torch

sys.path.insert(0, ".")
from tacdmpc_b200 import _lib  # noqa: E402
from tacdmpc_b200.ACMPC import ActorMPC  # noqa: E402
from tacdmpc_b200.dynamics import CartPole  # noqa: E402

dev = torch.device("cuda")


def timed(fn, reps=15, warm=4):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in ev:
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in ev)
    return ts[len(ts) // 2]


for T in (20, 15):
    for fused in (True, False):
        torch.manual_seed(0)
        f = CartPole()
        actor = ActorMPC(4, 1, T, 0.05, f_dyn=f, f_dyn_jac=f.jac, device="cuda", grad_method="analytic", fused=fused)
        for B in (512, 4096):
            obs = torch.rand(B, 4, device=dev) * 0.45 - 0.2          # U(-0.2, 0.25)^4, examples/02_demo_cartpole_AC.py:100
            acts = torch.randn(B, 1, device=dev)
            if B == 512:   # graph captures first (before any eager backward on the default stream)
                graphed = {b_: (actor.graphed(b_), actor.graphed_update(lambda lp, ent: -(lp.mean()) - 0.01 * ent.mean(),
                                                                         torch.zeros(b_, 4, device=dev), torch.zeros(b_, 1, device=dev)))
                           for b_ in (512, 4096)}

            def rollout_call():
                with torch.no_grad():
                    return actor(obs, deterministic=False)

            def update_step():
                actor.zero_grad(set_to_none=True)
                lp, ent = actor.evaluate_actions(obs, acts)
                (-(lp.mean()) - 0.01 * ent.mean()).backward()

            t_roll = timed(rollout_call)
            pol, upd = graphed[B]
            t_graph = timed(lambda: pol(obs))
            t_upd = timed(update_step)
            t_gupd = timed(lambda: upd(obs, acts))
            n0 = _lib.lib().tacd_kernel_launches()
            update_step()
            launches = _lib.lib().tacd_kernel_launches() - n0
            it = float(actor.mpc.iters_last.float().mean())
            print(f"T={T} {'compact-diagonal' if fused else 'dense           '} B={B:5d}: policy call {t_roll:7.3f} ms, graphed {t_graph:6.3f} ms ({B / t_graph / 1e3:7.3f} M obs/s)   "
                  f"PPO update step {t_upd:7.3f} ms, graphed {t_gupd:6.3f} ms ({B / t_gupd / 1e3:7.3f} M samples/s)   mean iters {it:5.2f}   library launches/step {launches}",
                  flush=True)
