"""The ActorMPC call site of the reference (ACMPC/actor.py) on top of the B200 solve.

Only the actor is mirrored: it is the caller of the hot path.  The critic, the PPO loop, the environment
manager and checkpointing are out of scope.
"""
from .actor import ActorMPC

__all__ = ["ActorMPC"]
