"""Device-timed PPO glue kernels next to the numpy port on the host (developer tool).

    python tools/bench_ppo.py
"""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from oracle import ppo_oracle as po  # noqa: E402  (CPU baseline only)
from tacdmpc_b200 import _abi, ops  # noqa: E402

dev = torch.device("cuda")


def timed(fn, reps=20, warm=5):
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
    return ts[len(ts) // 2] * 1e3  # us


def cpu_timed(fn, reps=3):
    best = 1e9
    for _ in range(reps):
        t = time.perf_counter()
        fn()
        best = min(best, time.perf_counter() - t)
    return best * 1e6


rng = np.random.default_rng(0)
for N, T in ((4096, 250), (65536, 250)):
    r, v = rng.standard_normal((N, T)).astype(np.float32), rng.standard_normal((N, T)).astype(np.float32)
    dr, dv = torch.from_numpy(r).to(dev), torch.from_numpy(v).to(dev)
    us = timed(lambda: ops.gae(dr, dv, 0.99, 0.97))
    cpu = cpu_timed(lambda: po.gae(r, v, 0.99, 0.97)) if N <= 4096 else float("nan")
    print(f"gae N={N} T={T}: {us:8.1f} us  ({16.0 * N * T / us / 1e3:7.1f} GB/s algorithmic)   numpy port {cpu:10.1f} us")
N, H = 512 * 8, 20
r, b = rng.standard_normal((N, H)).astype(np.float32), rng.standard_normal(N).astype(np.float32)
dr, db = torch.from_numpy(r).to(dev), torch.from_numpy(b).to(dev)
print(f"mpve N={N} H={H}: {timed(lambda: ops.discounted_targets(dr, db, 0.99)):8.1f} us   numpy port {cpu_timed(lambda: po.discounted_targets(r, b, 0.99)):8.1f} us")
for B in (4096, 65536, 1 << 20):
    spec = ops.Spec(nx=4, nu=1, T=1, dt=0.05, dyn_id=_abi.DYN_CARTPOLE, dyn_params=(1.0, 0.1, 0.5, 9.8))
    x = torch.rand(B, 4, device=dev) * 0.4 - 0.2
    a = torch.randn(B, 1, device=dev)
    st = torch.zeros(B, dtype=torch.int32, device=dev)
    rs = torch.rand(B, 4, device=dev) * 0.4 - 0.2
    us = timed(lambda: ops.env_step(spec, x, a, steps=st, reset_state=rs, max_steps=500, term_limit=(2.4, float("inf"), 0.2094, float("inf")),
                                    w_x=(10, 10, 100, 10), w_u=(0.01,)))
    xn, an, sn, rn = x.cpu().numpy(), a.cpu().numpy(), st.cpu().numpy(), rs.cpu().numpy()
    cpu = cpu_timed(lambda: po.env_step(lambda s, u: po.cartpole_f(s, u, 0.05, 1.0, 0.1, 0.5, 9.8), xn, an, sn, rn, max_steps=500,
                                        term_limit=(2.4, np.inf, 0.2094, np.inf), w_x=(10, 10, 100, 10), w_u=(0.01,))) if B <= 65536 else float("nan")
    print(f"env_step cartpole B={B}: {us:8.1f} us incl. host wrapper ({B / us:8.1f} M env-steps/s)   numpy port {cpu:10.1f} us")
