"""cProfile of the host side of one ActorMPC policy call / PPO update step (developer tool)."""
import cProfile
import pstats
import sys

import torch

sys.path.insert(0, ".")
from tacdmpc_b200.ACMPC import ActorMPC  # noqa: E402
from tacdmpc_b200.dynamics import CartPole  # noqa: E402

f = CartPole()
actor = ActorMPC(4, 1, 20, 0.05, f_dyn=f, f_dyn_jac=f.jac, device="cuda", grad_method="analytic")
obs = torch.rand(512, 4, device="cuda") * 0.45 - 0.2
acts = torch.randn(512, 1, device="cuda")


def policy():
    with torch.no_grad():
        actor(obs)


def update():
    actor.zero_grad(set_to_none=True)
    lp, ent = actor.evaluate_actions(obs, acts)
    (-(lp.mean()) - 0.01 * ent.mean()).backward()


for fn in (policy, update):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(200):
        fn()
    torch.cuda.synchronize()
    pr.disable()
    print("=====", fn.__name__)
    pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
