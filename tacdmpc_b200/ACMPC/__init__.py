"""The call sites of the reference's ACMPC package on top of the B200 solve.

``ActorMPC`` (ACMPC/actor.py) is the caller of the hot path; ``DeviceEnvManager`` (ACMPC/parallel_env.py) and
``compute_gae_and_returns`` / ``mpve_targets`` / ``collect_rollout`` (ACMPC/training_loop.py) are the batched pieces
around it.  ``ShardedPPOUpdate`` is the
mini-batch update of train() with the samples partitioned over ranks and one flat gradient all-reduce.  The critic
model, the outer loop and checkpointing are plain torch in the reference and out of scope.
"""
from .actor import ActorMPC, GraphedPolicy, GraphedUpdate
from .parallel_env import DeviceEnvManager, cartpole_envs, double_integrator_envs, uniform_reset
from .training_loop import ShardedPPOUpdate, collect_rollout, compute_gae_and_returns, mpve_targets

__all__ = ["ActorMPC", "GraphedPolicy", "GraphedUpdate", "DeviceEnvManager", "cartpole_envs", "double_integrator_envs", "uniform_reset",
           "collect_rollout", "compute_gae_and_returns", "mpve_targets", "ShardedPPOUpdate"]
