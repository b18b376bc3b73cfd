"""Host-side multi-GPU logic on CPU: world_size-2 gloo processes.

The solve shards over the batch with no data-path collective; what has to be right on the host is
 (1) the contiguous partition, (2) quirk Q4 — every rank's line search uses GLOBAL element 0's cost,
 broadcast from rank 0, (3) shared-parameter gradients are summed across ranks with one flat all-reduce.
The compute in these tests is done by the CPU oracle (test infrastructure); the product kernels are
covered by the -m gpu tests."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import parity as P


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, name, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import oracle as orc
        from tacdmpc_b200 import sharding
        from tacdmpc_b200.DifferentialMPC import GeneralQuadCost

        g = P.load(name)
        spec = P.spec_of(g)
        B = g["x0"].shape[0]
        lo, hi = sharding.shard_range(B, rank, world)
        sh = lambda a, batched: a[lo:hi] if batched else a
        Cb, Cfb = g["C"].ndim == 4, g["Cf"].ndim == 3

        # (2) quirk Q4 through the product helper: a stand-in controller holding this rank's shard
        class Ctl:
            compat = "reference"
            ls_cost_override = None
        ctl = Ctl()
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a))
        ctl.cost_module = GeneralQuadCost(spec["nx"], spec["nu"], t(sh(g["C"], Cb)), t(sh(g["c"], Cb)), t(sh(g["Cf"], Cfb)),
                                          t(sh(g["cf"], Cfb)))
        sharding.broadcast_line_search_cost(ctl)
        ls = None if ctl.ls_cost_override is None else tuple(x.numpy() for x in ctl.ls_cost_override)
        if Cb or Cfb:
            ref = P.ls_cost_of(g)
            for a, b in zip(ls, ref):
                assert np.array_equal(a, b), "rank %d does not hold global element 0's cost" % rank
        else:
            assert ls is None

        out = orc.forward(spec, g["x0"][lo:hi], sh(g["C"], Cb), sh(g["c"], Cb), sh(g["Cf"], Cfb), sh(g["cf"], Cfb),
                          g["xref"][lo:hi], g["uref"][lo:hi], g["Uinit"][lo:hi], ls_cost=ls)
        gU = np.zeros_like(out["U"])
        gU[:, 0] = 1
        bw = orc.backward(spec, out["X"], out["U"], sh(g["C"], Cb), g["xref"][lo:hi], g["uref"][lo:hi], out["tight"],
                          np.zeros_like(out["X"]), gU, Cf_batched=int(Cfb))
        # (3) shared-cost gradients: one flat all-reduce
        bw0 = {k: v.copy() for k, v in bw.items()}
        shared = [torch.from_numpy(bw[k]) for k, b in (("grad_C", Cb), ("grad_c", Cb), ("grad_Cf", Cfb), ("grad_cf", Cfb)) if not b]
        sharding.allreduce_sum_(shared)
        # the same sum through the flat bucket the bench / PPO step use (views handed out as result buffers, one all-reduce)
        names = [k for k, b in (("grad_C", Cb), ("grad_c", Cb), ("grad_Cf", Cfb), ("grad_cf", Cfb)) if not b]
        if names:
            bk = sharding.GradBucket({k: bw0[k].shape for k in names}, "cpu", torch.float64)
            for k in names:
                bk.views[k].copy_(torch.from_numpy(bw0[k]))
            bk.all_reduce_async()
            bk.join()
            for k, t in zip(names, shared):
                assert torch.equal(bk.views[k], t), k
        tmax = sharding.allreduce_max_scalar(float(rank + 1), "cpu")
        assert tmax == float(world)
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), lo=lo, hi=hi, X=out["X"], U=out["U"], iters=out["iters"],
                 **{k: v for k, v in bw.items()})
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("name", ["cartpole_actor_f64", "cartpole_shared_f64", "lti8_mixed_f64"])
def test_sharded_solve_equals_single_process(oracle, tmp_path, name):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), name, str(tmp_path)), nprocs=world, join=True)
    g = P.load(name)
    spec = P.spec_of(g)
    Cb, Cfb = g["C"].ndim == 4, g["Cf"].ndim == 3
    full = oracle.forward(spec, g["x0"], g["C"], g["c"], g["Cf"], g["cf"], g["xref"], g["uref"], g["Uinit"],
                          ls_cost=P.ls_cost_of(g))
    gU = np.zeros_like(full["U"])
    gU[:, 0] = 1
    fbw = oracle.backward(spec, full["X"], full["U"], g["C"], g["xref"], g["uref"], full["tight"], np.zeros_like(full["X"]), gU,
                          Cf_batched=int(Cfb))
    parts = [np.load(tmp_path / f"rank{r}.npz") for r in range(world)]
    assert parts[0]["lo"] == 0 and parts[-1]["hi"] == g["x0"].shape[0] and parts[0]["hi"] == parts[1]["lo"]
    # every batch element is an independent solve (quirk Q11): bit-identical to the single-process run
    assert np.array_equal(np.concatenate([p["X"] for p in parts]), full["X"])
    assert np.array_equal(np.concatenate([p["U"] for p in parts]), full["U"])
    assert np.array_equal(np.concatenate([p["iters"] for p in parts]), full["iters"])
    for k, batched in (("grad_C", Cb), ("grad_c", Cb), ("grad_Cf", Cfb), ("grad_cf", Cfb)):
        if batched:
            assert np.array_equal(np.concatenate([p[k] for p in parts]), fbw[k])
        else:   # summed across ranks by the all-reduce: same value on every rank, equal to the full-batch sum
            assert np.array_equal(parts[0][k], parts[1][k])
            np.testing.assert_allclose(parts[0][k], fbw[k], rtol=1e-12, atol=1e-12 * max(1.0, np.abs(fbw[k]).max()))


def _worker_diag(rank, world, port, name, out_dir):
    """with a DiagQuadCost the override must be element 0's DIAGONALS, and the controller must hand
    the kernels the layout of the path that runs (compact for C_diag, expanded for the dense fallback)."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from tacdmpc_b200 import sharding
        from tacdmpc_b200.DifferentialMPC import DiagQuadCost, DifferentiableMPCController
        from tacdmpc_b200.dynamics import CartPole
        g = P.load(name)
        B = g["x0"].shape[0]
        lo, hi = sharding.shard_range(B, rank, world)
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a))
        q, qf = np.diagonal(g["C"], axis1=-2, axis2=-1), np.diagonal(g["Cf"], axis1=-2, axis2=-1)
        cm = DiagQuadCost(4, 1, t(q[lo:hi]), t(g["c"][lo:hi]), t(qf[lo:hi]), t(g["cf"][lo:hi]), device="cpu")
        f = CartPole()
        ctl = DifferentiableMPCController(f, 1.0, 0.05, int(g["meta_T"]), cm, device="cpu", grad_method="analytic", f_dyn_jac=f.jac)
        sharding.broadcast_line_search_cost(ctl)
        ls = ctl.ls_cost_override
        assert ls is not None and ls[0].shape == q.shape[1:] and ls[2].shape == qf.shape[1:]
        for a, b in zip(ls, (q[0], g["c"][0], qf[0], g["cf"][0])):
            assert np.array_equal(a.numpy(), b), "rank %d does not hold global element 0's diagonals" % rank
        # what the two kernel families would be handed
        cd = ctl._ls_override(torch.float64, "cpu", True)
        dn = ctl._ls_override(torch.float64, "cpu", False)
        ref = P.ls_cost_of(g)
        assert all(np.array_equal(a.numpy(), b) for a, b in zip(dn, ref))
        assert np.array_equal(cd[0].numpy(), q[0]) and np.array_equal(cd[2].numpy(), qf[0])
        # a dense override (what round 1 installed) is reduced to its diagonal for the compact kernels
        ctl.ls_cost_override = tuple(t(a) for a in ref)
        cd2 = ctl._ls_override(torch.float64, "cpu", True)
        assert all(np.array_equal(a.numpy(), b.numpy()) for a, b in zip(cd, cd2))
        ctl.ls_cost_override = (t(ref[0][:, :2]), t(ref[1]), t(ref[2]), t(ref[3]))
        with pytest.raises(RuntimeError):
            ctl._ls_override(torch.float64, "cpu", True)
    finally:
        dist.destroy_process_group()


def test_diag_cost_line_search_broadcast(tmp_path):
    mp.spawn(_worker_diag, args=(2, _free_port(), "cartpole_actor_f64", str(tmp_path)), nprocs=2, join=True)


def test_shard_range_partitions_exactly():
    from tacdmpc_b200 import sharding
    for B in (0, 1, 7, 8, 65536, 65537):
        for world in (1, 2, 3, 8):
            spans = [sharding.shard_range(B, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
