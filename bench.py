#!/usr/bin/env python
"""bench.py — iLQR forward+backward solves/s (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--layout shared|per_element]

Workload (BASELINE.json configs[1]): cart-pole regulation, nx=4, nu=1, T=20, fp32,
analytic Jacobians, u in [-20, 20], B = 65536 solves PER GPU (weak scaling: the batch of independent
solves is partitioned across ranks, no data-path collective), synthetic seeded inputs.
One "step" = one forward solve + one implicit-differentiation backward over the batch, loss = U*[:,0].sum()
(what PPO differentiates), every input requiring a gradient.

`value`   : solves/s with inputs resident in HBM, CUDA-event timed on the launching stream, L2 flushed
            between timed steps, max over ranks.
`e2e`     : the same metric through the public API (DifferentiableMPCController + autograd): every step the
            states x0 and the warm start U_init are copied from pinned HOST memory (two-deep pipeline on a copy
            stream) and the actions + one cost gradient are read back; references are set once, as in the
            reference's regulation example.
`value`   : median of the per-step device times, max over ranks; mean / max / the list are in `timing`.
`roofline`: the dominant kernel (the forward solve) against its BINDING roof (max of the bytes term and the
            flops term): the FP32 pipe (model flops of the measured mean iteration count / launch time vs the nominal
            FP32 peak, with an in-run FMA microbenchmark beside it); the HBM view is the sub-object `hbm_view`.
`extra_configs`: BASELINE configs 2b, 3, 4, 5 at their sizes (N=1): solves/s, forward / backward ms, mean iterations.
`cpu_baseline`: the CPU oracle ("port": the reference is pure Python and does not travel to the GPU box)
            on all host threads on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

NX, NU, T, DT = 4, 1, 20, 0.05
N = NX + NU
B_PER_GPU = 65536
METRIC = "iLQR fwd+bwd solves/sec"
UNIT = "solves/s"


# ------------------------------------------------------------------------------------------- workload
def make_inputs(B, layout, seed, torch):
    """Q=diag(10,1,100,1), R=0.1, C_final[:4,:4]=10Q, refs 0,
    x0=[0,0,0.2,0]+N(0,1)*[0.5,0.5,0.2,0.2], U_init=0.  layout 'per_element' scales every element's cost
    by U(0.5,1.5) (what ActorMPC-style per-element costs look like to the kernel)."""
    g = torch.Generator().manual_seed(seed)
    C = torch.zeros(T, N, N)
    C[:, :NX, :NX] = torch.diag(torch.tensor([10.0, 1.0, 100.0, 1.0]))
    C[:, NX:, NX:] = 0.1
    c = torch.zeros(T, N)
    Cf = torch.zeros(N, N)
    Cf[:NX, :NX] = 10 * C[0, :NX, :NX]
    cf = torch.zeros(N)
    x0 = torch.tensor([0.0, 0.0, 0.2, 0.0]) + torch.randn(B, NX, generator=g) * torch.tensor([0.5, 0.5, 0.2, 0.2])
    if layout in ("per_element", "diag"):
        s = 0.5 + torch.rand(B, 1, 1, 1, generator=g)
        C = (C.unsqueeze(0) * s).contiguous()
        c = c.unsqueeze(0).expand(B, T, N).contiguous()
        Cf = (Cf.unsqueeze(0) * s[:, 0]).contiguous()
        cf = cf.unsqueeze(0).expand(B, N).contiguous()
        if layout == "diag":   # the same per-element costs as compact diagonals (ActorMPC layout)
            C = torch.diagonal(C, dim1=-2, dim2=-1).contiguous()
            Cf = torch.diagonal(Cf, dim1=-2, dim2=-1).contiguous()
    return dict(x0=x0, C=C, c=c, Cf=Cf, cf=cf, xref=torch.zeros(B, T + 1, NX), uref=torch.zeros(B, T, NU),
                Uinit=torch.zeros(B, T, NU))


def algorithmic_bytes(layout):
    """fp32 bytes per solve, inputs read once and outputs written once, A/B recomputed."""
    xu = (T + 1) * NX + T * NU
    cost = T * N * N + T * N + N * N + N
    fwd = 4 * (NX + xu + T * NU + T * NU) + 4 * xu + T * NU          # x0, xref+uref, U_init | X*, U*, tight
    bwd_r = 4 * (3 * xu) + T * NU                                    # X*,U*, refs, gX,gU | tight
    bwd_w = 4 * (xu + NX)                                            # grad_xref, grad_uref, grad_x0
    if layout == "per_element":
        fwd += 4 * cost
        bwd_r += 4 * T * N * N
        bwd_w += 4 * cost
    if layout == "diag":
        fwd += 4 * (2 * T * N + 2 * N)
        bwd_r += 4 * T * N
        bwd_w += 4 * (2 * T * N + 2 * N)
    return fwd, bwd_r + bwd_w


def flops_per_solve(mean_iters):
    """SURVEY §8d flop model (FMA = 2) for cart-pole: F_dyn=45, F_jac=90."""
    nx, nu, n = NX, NU, N
    F_dyn, F_jac = 45, 90
    F_quad = T * (2 * n * n + n) + 2 * n * n + n
    F_ric = T * (4 * nx ** 3 + 10 * nx * nx * nu + 4 * nx * nu * nu + 2 * nx * nx + 8 * nx * nu + nu ** 3 / 3 + 4 * nu * nu + nu)
    F_ls = 5 * (T * (2 * nx * nu + 4 * nu + nx + F_dyn) + T * (2 * n * n + 3 * n) + 2 * nx * nx + 3 * nx)
    F_acc = T * (2 * n * n + 3 * n) + 2 * nx * nx + 3 * nx
    F_iter = F_quad + T * F_jac + F_ric + F_ls + F_acc
    F_final = F_quad + T * F_jac
    F_bwd = F_ric + T * (2 * nx * nx + 6 * nx * nu) + 3 * T * n * n
    return mean_iters * F_iter + F_final, F_bwd


# ------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """SM clock and throttle reasons DURING the timed region (B200_PROFILING.md recipe).  The timed region is a
    few tens of milliseconds, shorter than nvidia-smi's polling period, so NVML is polled directly from a thread
    (~1 kHz); `nvidia-smi -lms` is the fallback when NVML cannot be loaded."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index, uuid=None):
        self.rows, self.proc, self.index, self.uuid = [], None, index, uuid
        self.samples, self.max_mhz, self.reasons, self._stop, self.th, self.nvml = [], None, set(), False, None, None
        self.armed = False

    def sample(self):
        """One NVML reading (also called from the timed loop itself, while the GPU works through the queued steps: the
        polling thread alone can be starved by the interpreter lock for most of a 16 ms region)."""
        n, h = self.nvml, getattr(self, "handle", None)
        if n is None or h is None:
            return
        bits = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20}
        try:
            self.samples.append(float(n.nvmlDeviceGetClockInfo(h, n.NVML_CLOCK_SM)))
            r = int(n.nvmlDeviceGetCurrentClocksEventReasons(h)) if hasattr(n, "nvmlDeviceGetCurrentClocksEventReasons") \
                else int(n.nvmlDeviceGetCurrentClocksThrottleReasons(h))
            for name, b in bits.items():
                if r & b:
                    self.reasons.add(name)
        except Exception:
            pass

    def _nvml_loop(self):
        # samples are taken only while a timed region is open (`armed`), from this thread only: round 1 also read NVML from
        # the launching thread inside the timed loop, which stalled the launches of that rank
        while not self._stop:
            if self.armed:
                self.sample()
            time.sleep(0.002)

    def __enter__(self):
        try:
            import pynvml as n
            n.nvmlInit()
            h = None
            if self.uuid:
                for u in (self.uuid, "GPU-" + self.uuid):
                    try:
                        h = n.nvmlDeviceGetHandleByUUID(u.encode() if isinstance(u, str) else u)
                        break
                    except Exception:
                        h = None
            if h is None:
                h = n.nvmlDeviceGetHandleByIndex(self.index)
            self.nvml, self.handle = n, h
            self.max_mhz = float(n.nvmlDeviceGetMaxClockInfo(h, n.NVML_CLOCK_SM))
            self.th = threading.Thread(target=self._nvml_loop, daemon=True)
            self.th.start()
            return self
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def __exit__(self, *a):
        self._stop = True
        if self.nvml is not None:
            if self.th is not None:
                self.th.join(timeout=1)
            return
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        if self.nvml is not None:
            sm = sorted(self.samples)
            if not sm:
                return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0, "source": "nvml"}
            return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(sm),
                    "source": "nvml"}
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "source": "nvidia-smi"}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi"}


# ------------------------------------------------------------------------------------------- CPU arm
def cpu_port_run(layout, seconds_target, threads=None):
    """Forward+backward of the same workload on the CPU oracle (kind 'port'), all host threads, on a
    bounded sample sized to ~seconds_target of CPU work.  Returns (solves/s, cores, sample description)."""
    import numpy as np
    import torch
    from oracle import oracle as orc
    # all host threads this process may use; torchrun exports OMP_NUM_THREADS=1, which must not apply here
    orc.set_num_threads(threads or len(os.sched_getaffinity(0)))
    cores = orc.num_threads()
    spec = dict(nx=NX, nu=NU, T=T, dt=DT, dyn="cartpole", dyn_params=(1.0, 0.1, 0.5, 9.8), reg_eps=1e-6, max_iter=40,
                u_min=[-20.0], u_max=[20.0])

    def run(Bs):
        # the oracle restates the reference, which only knows dense costs: 'diag' is timed as 'per_element'
        inp = {k: v.numpy() for k, v in make_inputs(Bs, "per_element" if layout == "diag" else layout, 0, torch).items()}
        ls = None
        if layout != "shared":
            ls = (inp["C"][0], inp["c"][0], inp["Cf"][0], inp["cf"][0])
        t0 = time.perf_counter()
        f = orc.forward(spec, inp["x0"], inp["C"], inp["c"], inp["Cf"], inp["cf"], inp["xref"], inp["uref"], inp["Uinit"],
                        ls_cost=ls, want_AB=False, want_trace=False)
        gU = np.zeros_like(f["U"]); gU[:, 0] = 1
        orc.backward(spec, f["X"], f["U"], inp["C"], inp["xref"], inp["uref"], f["tight"], np.zeros_like(f["X"]), gU,
                     Cf_batched=int(layout != "shared"),
                     want=("grad_C", "grad_c", "grad_Cf", "grad_cf", "grad_x0", "grad_xref", "grad_uref"))
        return time.perf_counter() - t0, float(f["iters"].mean())

    t_probe, _ = run(2048)
    Bs = int(min(16 * B_PER_GPU, max(2048, 2048 * seconds_target / max(t_probe, 1e-6))))
    t, mi = run(Bs)
    return Bs / t, cores, f"B={Bs} of the same seeded workload, fwd+bwd, {t:.1f} s, mean iters {mi:.2f}", Bs, t


# ------------------------------------------------------------------------------------------- main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--layout", default="shared", choices=["shared", "per_element", "diag"])
    ap.add_argument("--batch", type=int, default=B_PER_GPU, help="solves per GPU")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the extra_configs table (BASELINE configs 2b-5, N=1 only)")
    args = ap.parse_args()
    # Contract: stdout carries ONE JSON line.  Libraries write there too (NCCL prints "NCCL version ..." on the first
    # communicator), so file descriptor 1 is pointed at stderr for the run and the line goes out through a saved copy.
    sys.stdout.flush()
    out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    W = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    config = {"workload": f"cartpole regulation iLQR nx=4 nu=1 T=20 analytic Jacobians, B={args.batch} per GPU "
                          f"(BASELINE configs[1] / SURVEY 8d config 2{dict(shared='a shared C', per_element='b per-element C', diag='b per-element diagonal C in compact form (8f N1)')[args.layout]})",
              "cost_layout": args.layout, "B_per_gpu": args.batch, "T": T, "nx": NX, "nu": NU, "max_iter": 40,
              "loss": "U*[:,0].sum(), all 7 inputs require grad", "sharding": f"batch x{world}, no data-path collective",
              "l2": "flushed (256 MiB write) between timed steps"}

    # ---------------------------------------------------------------- reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return 0
        vals = []
        for i in range(args.warmup + args.steps):
            v, cores, sample, Bs, t = cpu_port_run(args.layout, max(2.0, args.cpu_seconds / max(args.steps, 1)))
            if i >= args.warmup:
                vals.append((v, t, Bs))
        value = sum(b for _, _, b in vals) / sum(t for _, t, _ in vals)
        line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": 1e3 * sum(t for _, t, _ in vals) / len(vals), "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample + " per step"},
                "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0,
                "note": "CPU oracle (C restatement of the reference, OpenMP over the batch); the reference itself is "
                        "pure Python and cannot travel to the GPU box (DESIGN.md)"}
        print(json.dumps(line), file=out, flush=True)
        return 0

    # ---------------------------------------------------------------- B200 arm
    import torch
    import torch.distributed as dist
    from tacdmpc_b200 import _lib, ops, sharding
    from tacdmpc_b200 import workloads as WL
    from tacdmpc_b200.DifferentialMPC import DifferentiableMPCController, GeneralQuadCost
    from tacdmpc_b200.dynamics import CartPole

    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (there is no CPU fallback)")
    _lib.lib()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B = args.batch
    host = make_inputs(B, args.layout, 1000 + rank, torch)
    d = {k: v.to(dev) for k, v in host.items()}
    f = CartPole()
    spec = ops.Spec(nx=NX, nu=NU, T=T, dt=DT, dyn_id=f.dyn_id, dyn_params=f.params, reg_eps=1e-6, max_iter=40,
                    u_min=[-20.0], u_max=[20.0])
    ls = None
    diag = args.layout == "diag"
    if args.layout != "shared":
        # quirk Q4: every rank's line search uses GLOBAL element 0's cost
        ls0 = torch.cat([d["C"][0].reshape(-1), d["c"][0].reshape(-1), d["Cf"][0].reshape(-1), d["cf"][0].reshape(-1)])
        if world > 1:
            dist.broadcast(ls0, src=0)
        o = 0
        ls = []
        for shp in (((T, N), (T, N), (N,), (N,)) if diag else ((T, N, N), (T, N), (N, N), (N,))):
            k = 1
            for s_ in shp:
                k *= s_
            ls.append(ls0[o:o + k].view(shp).contiguous())
            o += k
        ls = tuple(ls)
    gX = torch.zeros(B, T + 1, NX, device=dev)
    gU = torch.zeros(B, T, NU, device=dev)
    gU[:, 0] = 1.0
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    want = ("grad_C", "grad_c", "grad_Cf", "grad_cf", "grad_x0", "grad_xref", "grad_uref")
    shared = args.layout == "shared"
    # Shared-cost gradients are sums over the whole (sharded) batch: the backward kernels write them straight into one flat
    # bucket (no gather copy), and the cross-rank sum is ONE all-reduce on a communication stream that the next step's solve
    # does not wait for (two buckets, alternating: a consumer has a full step to read the reduced gradient).  Round 1 ran
    # cat + all-reduce + four copies on the launching stream inside every step's window.
    buckets = [sharding.GradBucket({"grad_C": (T, N, N), "grad_c": (T, N), "grad_Cf": (N, N), "grad_cf": (N,)}, dev) for _ in range(2)] \
        if shared else None

    def fwd_call():
        return ops.ilqr_forward(spec, d["x0"], d["C"], d["c"], d["Cf"], d["cf"], d["xref"], d["uref"], d["Uinit"], ls_cost=ls,
                                want_trace=False, C_diag=diag)

    def bwd_call(fw, slot):
        return ops.ilqr_backward(spec, fw["X"], fw["U"], d["C"], d["xref"], d["uref"], fw["tight"], gX, gU,
                                 Cf_batched=not shared, want=want, C_diag=diag, out=buckets[slot].views if shared else None)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(W):                       # eager warm-up (also what CUDA-graph capture needs to have happened)
        fw = fwd_call()
        bw = bwd_call(fw, i & 1)
        if shared:
            buckets[i & 1].all_reduce_async()
        flush.fill_(1)
    if shared:
        for bk in buckets:
            bk.join()
    barrier()
    # The device-timed loop replays the forward and the backward from CUDA graphs (one forward graph, one backward graph
    # per bucket): the host then spends ~20 us per step instead of ~300 us in the Python / ctypes launch path, so with 8
    # ranks on one box no rank's GPU ever waits for its host (the solve never synchronises the host and allocates only
    # through torch, so it is capturable as it is).  TACD_BENCH_NO_GRAPH=1 times the eager calls instead.
    g_fwd = g_bwd = None
    if not os.environ.get("TACD_BENCH_NO_GRAPH"):
        try:
            g_fwd = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g_fwd, capture_error_mode="thread_local"):
                fw = fwd_call()
            g_bwd = []
            for sl in range(2 if shared else 1):
                g_ = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g_, capture_error_mode="thread_local"):
                    bw = bwd_call(fw, sl)
                g_bwd.append(g_)
        except Exception as ex:   # noqa: BLE001
            print(f"[bench] CUDA-graph capture of the device-timed step failed ({type(ex).__name__}: {ex}); eager loop", file=sys.stderr)
            g_fwd = g_bwd = None
            torch.cuda.synchronize(dev)
    launches_per_step = None
    l0 = int(_lib.lib().tacd_kernel_launches())
    fw_e = fwd_call()
    bwd_call(fw_e, 0)
    launches_per_step = int(_lib.lib().tacd_kernel_launches()) - l0     # kernels of this library per step (graph replays re-issue them)
    barrier()
    K = args.steps
    ev = [tuple(torch.cuda.Event(enable_timing=True) for _ in range(3)) for _ in range(K)]
    try:
        dev_uuid = str(torch.cuda.get_device_properties(dev).uuid)
    except Exception:
        dev_uuid = None
    clk = ClockSampler(local_rank, dev_uuid)
    clk.__enter__()
    barrier()
    clk.armed = True
    t_wall0 = time.perf_counter()
    for i in range(K):
        flush.fill_(i & 0xFF)          # L2 flush, outside the timed events
        sl = i & 1
        if shared:
            buckets[sl].join()         # the all-reduce that last used this bucket (two steps ago) is long done: no stall
        ev[i][0].record()
        if g_fwd is not None:
            g_fwd.replay()
        else:
            fw = fwd_call()
        ev[i][1].record()
        if g_bwd is not None:
            g_bwd[sl if shared else 0].replay()
        else:
            bw = bwd_call(fw, sl)
        ev[i][2].record()
        if shared:
            buckets[sl].all_reduce_async()     # communication stream; `done` event marks the reduced gradient of step i
    if shared:
        for bk in buckets:
            bk.join()
    barrier()                                  # every rank's solves AND all-reduces are complete inside the bracket
    t_wall = time.perf_counter() - t_wall0
    clk.armed = False
    gpu_launches = launches_per_step * K
    mean_iters = float(fw["iters"].float().mean().item())
    step_ms = [e[0].elapsed_time(e[2]) for e in ev]
    fwd_ms_l = [e[0].elapsed_time(e[1]) for e in ev]
    med = lambda v: sorted(v)[len(v) // 2]
    # median of the per-step device times (a straggler step shows in max / in the list, not in the headline)
    t_step = sharding.allreduce_max_scalar(med(step_ms), dev)
    t_fwd_step = sharding.allreduce_max_scalar(med(fwd_ms_l), dev)
    t_step_mean = sharding.allreduce_max_scalar(sum(step_ms) / K, dev)
    t_step_max = sharding.allreduce_max_scalar(max(step_ms), dev)
    t_bracket = sharding.allreduce_max_scalar(t_wall, dev)
    value = world * B / (t_step * 1e-3)
    t_tot_max, t_fwd_max = t_step * 1e-3 * K, t_fwd_step * 1e-3 * K

    # ---------------------------------------------------------------- e2e through the public API, host buffers
    pin = {k: v.pin_memory() for k, v in host.items()}
    Cd = d["C"].clone().requires_grad_(True)
    if diag:
        from tacdmpc_b200.DifferentialMPC import DiagQuadCost
        cm = DiagQuadCost(NX, NU, Cd, d["c"], d["Cf"], d["cf"], device=str(dev), x_ref=d["xref"], u_ref=d["uref"])
    else:
        cm = GeneralQuadCost(NX, NU, Cd, d["c"], d["Cf"], d["cf"], device=str(dev), x_ref=d["xref"], u_ref=d["uref"])
    mpc = DifferentiableMPCController(f, T * DT, DT, T, cm, u_min=torch.tensor([-20.0]), u_max=torch.tensor([20.0]),
                                      device=str(dev), grad_method="analytic", f_dyn_jac=f.jac)
    if ls is not None:
        mpc.ls_cost_override = ls
    act_host = torch.empty(B, NU).pin_memory()
    gC_host = torch.empty(Cd.shape if args.layout == "shared" else Cd.shape[1:]).pin_memory()
    h2d = sum(pin[k].numel() * 4 for k in ("x0", "Uinit"))
    d2h = act_host.numel() * 4 + gC_host.numel() * 4

    # Two-deep pipeline: the host->device copy of step i+1's inputs runs on a copy stream while step i
    # computes; every step's copies (and the read-back of its result) are inside the timed region.
    copy_s = torch.cuda.Stream(device=dev)
    comp_s = torch.cuda.current_stream(dev)
    # Per-step host inputs: the states x0 (what the environments produce) and the warm start U_init.  The
    # references of this regulation workload are set once with set_reference(), as the reference's example
    # does (examples/02_demo_cartpole_regulation.py:150-154); they stay device tensors that require grad.
    names = ("x0", "Uinit")
    xr_dev = d["xref"].clone().requires_grad_(True)
    ur_dev = d["uref"].clone().requires_grad_(True)
    cm.set_reference(xr_dev, ur_dev)
    slots = [{k: torch.empty_like(d[k]) for k in names} for _ in range(2)]
    ev_up = [torch.cuda.Event() for _ in range(2)]
    ev_free = [torch.cuda.Event() for _ in range(2)]

    def upload(slot):
        with torch.cuda.stream(copy_s):
            copy_s.wait_event(ev_free[slot])          # the step that last used this slot has finished
            with torch.no_grad():                     # (x0 is a leaf that requires grad)
                for k in names:
                    slots[slot][k].copy_(pin[k], non_blocking=True)
            ev_up[slot].record(copy_s)

    def step_core(slot):
        """The device part of one public-API step on the inputs of `slot`: solve, loss, backward, read-back of the
        actions.  Returns the gradient that is read back (this rank's contribution when sharded)."""
        x0 = slots[slot]["x0"]
        Cd.grad = xr_dev.grad = ur_dev.grad = x0.grad = None
        X, U = mpc(x0, slots[slot]["Uinit"])
        U[:, 0].sum().backward()
        act_host.copy_(U.detach()[:, 0], non_blocking=True)
        return Cd.grad if args.layout == "shared" else Cd.grad[0]   # per-element: grads stay on the device, read one back

    tail_s = torch.cuda.Stream(device=dev)
    tail_ready = [torch.cuda.Event() for _ in range(2)]
    tail_done = [torch.cuda.Event() for _ in range(2)]

    def step_tail(gC, slot):
        """Cross-rank sum of the shared-cost gradient (the path's only collective) and its read-back, on a side stream: the
        next step's solve does not wait for them (they are joined before the slot's buffers are reused and before the
        closing barrier, so every step's collective and read-back are inside the timed region)."""
        tail_ready[slot].record(comp_s)
        with torch.cuda.stream(tail_s):
            tail_s.wait_event(tail_ready[slot])
            gC.record_stream(tail_s)
            if world > 1 and args.layout == "shared":
                dist.all_reduce(gC, op=dist.ReduceOp.SUM)
            gC_host.copy_(gC, non_blocking=True)
            tail_done[slot].record(tail_s)

    def step_body(slot):
        step_tail(step_core(slot), slot)

    for sl in slots:
        sl["x0"].requires_grad_(True)
    # The solve never synchronises the host and allocates only through torch's allocator, so a whole step
    # (forward, autograd backward, read-back) is CUDA-graph capturable: one graph per input slot removes the
    # ~0.2 ms of Python / launch gaps per step that the eager loop leaves between kernels.  The NCCL all-reduce of the
    # sharded case and the gradient read-back stay outside the graph (step_tail); falls back to eager if capture fails.
    graphs = None
    g_static = [None, None]
    if not os.environ.get("TACD_BENCH_NO_GRAPH"):
        try:
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(comp_s)
            with torch.cuda.stream(side):
                for sl in range(2):
                    with torch.no_grad():
                        for k in names:
                            slots[sl][k].copy_(pin[k], non_blocking=True)
                    for _ in range(2):
                        step_body(sl)
            comp_s.wait_stream(side)
            torch.cuda.synchronize(dev)
            graphs = []
            for sl in range(2):
                g_ = torch.cuda.CUDAGraph()
                Cd.grad = xr_dev.grad = ur_dev.grad = slots[sl]["x0"].grad = None
                # (thread_local: the NCCL watchdog thread of a sharded run may query events while this thread captures)
                with torch.cuda.graph(g_, capture_error_mode="thread_local"):
                    g_static[sl] = step_core(sl)
                graphs.append(g_)
        except Exception as ex:   # noqa: BLE001
            print(f"[bench] CUDA-graph capture of the e2e step failed ({type(ex).__name__}: {ex}); eager loop", file=sys.stderr)
            if os.environ.get("TACD_BENCH_DEBUG"):
                import traceback
                traceback.print_exc()
            graphs = None
            torch.cuda.synchronize(dev)

    def run_e2e(n):
        for e_ in ev_free:
            e_.record(comp_s)
        for e_ in tail_done:
            e_.record(tail_s)
        upload(0)
        for i in range(n):
            slot = i & 1
            if i + 1 < n:
                upload(slot ^ 1)
            comp_s.wait_event(ev_up[slot])
            comp_s.wait_event(tail_done[slot])     # the tail that last read this slot's gradient (two steps ago)
            if graphs is not None:
                graphs[slot].replay()
                step_tail(g_static[slot], slot)
            else:
                step_body(slot)
            ev_free[slot].record(comp_s)
        for e_ in tail_done:
            comp_s.wait_event(e_)

    run_e2e(W)
    barrier()
    clk.armed = True
    t0 = time.perf_counter()
    run_e2e(args.steps)
    barrier()
    t_e2e = sharding.allreduce_max_scalar(time.perf_counter() - t0, dev)
    clk.armed = False
    clk.__exit__(None, None, None)
    clocks = clk.summary()
    e2e_value = world * B * args.steps / t_e2e
    # the read-back of the last e2e step must be the actions of the device-timed solve (same inputs)
    e2e_checked = bool(torch.equal(act_host, fw["U"][:, 0].cpu()))
    if not e2e_checked:
        raise SystemExit("bench.py: the e2e path returned different actions than the device-timed solve")

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---------------------------------------------------------------- roofline + CPU baseline
    peaks = {}
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peaks = json.load(open(pk))
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    fwd_b, bwd_b = algorithmic_bytes(args.layout)
    fwd_ms = t_fwd_step
    bwd_ms = t_step - t_fwd_step
    ach = fwd_b * B / (fwd_ms * 1e-3) / 1e9
    ach_b = bwd_b * B / (bwd_ms * 1e-3) / 1e9
    F_fwd, F_bwd = flops_per_solve(mean_iters)
    sm_mhz = clocks.get("sm_mhz") or 1965.0
    fp32_peak_nominal = 148 * 128 * 2 * 1.965e9 / 1e12
    fp32_peak_at_clock = 148 * 128 * 2 * sm_mhz * 1e6 / 1e12
    # FP32-pipe microbenchmark on this GPU, this run: dependent-FMA chains, scalar FFMA and packed FFMA2
    probe = {}
    try:
        for nm, kind in (("ffma", 0), ("ffma2", 1)):
            tf, ms_ = ops.fp32_probe(kind, iters=8192, device=dev)
            probe[nm + "_tflops"] = tf
        probe["how"] = "tacd_fp32_probe: 148*8 CTAs x 256 threads x 16 chains x 8192 dependent FMAs, best of 5, CUDA events"
    except Exception as ex:   # noqa: BLE001
        probe = {"error": f"{type(ex).__name__}: {ex}"}
    fp32_measured = max(probe.get("ffma_tflops", 0.0), probe.get("ffma2_tflops", 0.0)) or None
    traffic = limiter = None
    tj = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tj):
        try:
            facts = json.load(open(tj))
            traffic = facts.get(f"forward_{args.layout}")
            limiter = facts.get(f"forward_{args.layout}_limiter")
        except Exception:
            traffic = None
    fwd_kernel = "ilqr_forward_small<float,4,1,cartpole>" if os.environ.get("TACD_FWD_KERNEL") == "thread" else \
        "ilqr_init_rollout + ilqr_forward_group<float,4,1,cartpole> (five lanes per element)"
    ach_tf = F_fwd * B / (fwd_ms * 1e-3) / 1e12
    hbm_view = {"bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak, "traffic": traffic,
                "peak_source": peak_src, "algorithmic_bytes_per_solve": fwd_b}
    fp32_view = {"bound": "fp32", "achieved": ach_tf, "peak": fp32_peak_nominal, "unit": "TFLOP/s", "frac": ach_tf / fp32_peak_nominal,
                 "peak_source": "nominal 148 SM x 128 lanes x 2 x 1.965 GHz (no driver-measured FP32 figure in MEASURED_PEAKS.json); "
                                "the in-run FMA probe is given beside it",
                 "flops_per_solve": F_fwd, "peak_tflops_at_measured_clock": fp32_peak_at_clock, "probe": probe,
                 "frac_of_measured_probe": (ach_tf / fp32_measured) if fp32_measured else None}
    # roofline.achieved = max(bytes term, flops term) — the BINDING roof leads, the other view sits beside it
    # (the contract's enum offers hbm|tensor; the forward solve is bound by the FP32 / issue pipes, so the bound is named "fp32")
    lead, other = (fp32_view, hbm_view) if fp32_view["frac"] >= hbm_view["frac"] else (hbm_view, fp32_view)
    roofline = dict(lead)
    roofline.update({"kernel": fwd_kernel, "launch_ms": fwd_ms, "mean_iters": mean_iters, "limiter": limiter,
                     "traffic": traffic, ("hbm_view" if lead is fp32_view else "fp32_view"): other,
                     "backward": {"kernel": "ilqr_backward_small<float,4,1,cartpole>" + (" + reduce_shared_grads" if shared else ""),
                                  "bound": "hbm", "achieved": ach_b, "peak": hbm_peak,
                                  "unit": "GB/s", "frac": ach_b / hbm_peak, "algorithmic_bytes_per_solve": bwd_b, "launch_ms": bwd_ms}})
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        v, cores, sample, _, _ = cpu_port_run(args.layout, args.cpu_seconds)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
               "reference_python": "the unmodified reference (pure Python, does not travel to this box) measured 141 solves/s on 8 cores "
                                   "in the build container at B=1024 (SURVEY 6)"}
    extra = None
    if world == 1 and not args.no_extra:
        try:
            extra = extra_configs(torch, ops, WL, dev, flush)
        except Exception as ex:   # noqa: BLE001
            extra = [{"error": f"{type(ex).__name__}: {ex}"}]
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": W,
            "ms_per_step": t_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": config, "clocks": clocks,
            "timing": {"statistic": "median over the K steps of each step's CUDA-event time (forward + backward), max over ranks",
                       "ms_per_step_mean": t_step_mean, "ms_per_step_max": t_step_max, "fwd_ms": fwd_ms, "bwd_ms": bwd_ms,
                       "bracket_ms_per_step": 1e3 * t_bracket / K, "bracket": "barrier + synchronize on both sides; includes the L2 "
                       "flushes, the host loop and, sharded, every step's all-reduce (overlapped on a communication stream)",
                       "mode": "cuda_graph" if g_fwd is not None else "eager", "step_ms_rank0": [round(v, 4) for v in step_ms]},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": 1e3 * t_e2e / args.steps, "mode": "cuda_graph" if graphs is not None else "eager",
                    "checked_against_device_solve": e2e_checked},
            "gpu_launches": gpu_launches, "roofline": roofline, "cpu_baseline": cpu, "extra_configs": extra}
    print(json.dumps(line), file=out, flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


# ------------------------------------------------------------------------------------------- the other configurations
def extra_configs(torch, ops, WL, dev, flush, reps=10, warm=3):
    """BASELINE.json configs 2b, 3, 4, 5 at their sizes on one GPU: device-timed forward / backward (median of `reps`,
    L2 flushed between), solves/s, mean iterations, fraction of the binding roof.  Same seeded workloads as the BASELINE-size
    parity tests (tacdmpc_b200/workloads.py)."""
    def timed(fn):
        for _ in range(warm):
            o = fn()
        torch.cuda.synchronize(dev)
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
        for a, b in ev:
            flush.fill_(1)
            a.record()
            o = fn()
            b.record()
        torch.cuda.synchronize(dev)
        ts = sorted(a.elapsed_time(b) for a, b in ev)
        return ts[len(ts) // 2], o

    cases = [("2a cartpole shared C, B=16384", lambda: WL.cartpole(16384, "shared"), "shared"),
             ("2a cartpole shared C, B=4096", lambda: WL.cartpole(4096, "shared"), "shared"),
             ("2a cartpole shared C, B=1024", lambda: WL.cartpole(1024, "shared"), "shared"),
             ("2b cartpole per-element dense C, B=65536", lambda: WL.cartpole(65536, "per_element"), "per_element"),
             ("2b cartpole per-element diagonal C (compact), B=65536", lambda: WL.cartpole(65536, "diag"), "diag"),
             ("3(i) unicycle T=30 loose box, B=16384", lambda: WL.unicycle(16384, tight=False), "shared"),
             ("3(ii) unicycle T=30 tight box, B=16384", lambda: WL.unicycle(16384, tight=True), "shared"),
             ("4 ActorMPC-shaped diagonal costs, B=512", lambda: WL.ppo_actor(512), "diag"),
             ("4 ActorMPC-shaped diagonal costs, B=4096", lambda: WL.ppo_actor(4096), "diag"),
             ("5 LTI nx=8 nu=4 T=50, B=4096", lambda: WL.lti(8, 4), "per_element"),
             ("5 LTI nx=16 nu=4 T=50, B=4096", lambda: WL.lti(16, 4), "per_element"),
             ("5 LTI nx=16 nu=8 T=50, B=4096", lambda: WL.lti(16, 8), "per_element"),
             ("5 LTI nx=32 nu=8 T=50, B=4096", lambda: WL.lti(32, 8), "per_element"),
             ("5 LTI nx=32 nu=8 T=50 spectral radius 0.9, B=4096", lambda: WL.lti(32, 8, rho=0.9), "per_element"),
             ("5 LTI nx=8 nu=4 T=50 spectral radius 0.9, B=4096", lambda: WL.lti(8, 4, rho=0.9), "per_element")]
    peaks = {}
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peaks = json.load(open(pk))
    hbm = float(peaks.get("hbm_gbs", 6650.0)) * 1e9
    fp32 = 148 * 128 * 2 * 1.965e9
    rows = []
    for name, make, layout in cases:
        w = make()
        s = WL.gpu_spec(w["spec"], dev)
        d = {k: v.to(dev) for k, v in w["inp"].items()}
        ls = None if w["ls"] is None else tuple(t.to(dev) for t in w["ls"])
        Bc, Tc = d["x0"].shape[0], s.T
        t_f, fw = timed(lambda: ops.ilqr_forward(s, d["x0"], d["C"], d["c"], d["Cf"], d["cf"], d["xref"], d["uref"], d["Uinit"], ls_cost=ls,
                                                 want_trace=False, C_diag=w["C_diag"]))
        gX = torch.zeros(Bc, Tc + 1, s.nx, device=dev)
        gU = torch.zeros(Bc, Tc, s.nu, device=dev)
        gU[:, 0] = 1
        t_b, _ = timed(lambda: ops.ilqr_backward(s, fw["X"], fw["U"], d["C"], d["xref"], d["uref"], fw["tight"], gX, gU,
                                                 Cf_batched=w["Cf_batched"], C_diag=w["C_diag"]))
        it = fw["iters"].float()
        mi = float(it.mean())
        bf, bb = WL.bytes_per_solve(w["spec"], layout)
        ff, fb = WL.flops_per_solve(w["spec"], mi)
        fr_f = max(bf * Bc / (t_f * 1e-3) / hbm, ff * Bc / (t_f * 1e-3) / fp32)
        fr_b = max(bb * Bc / (t_b * 1e-3) / hbm, fb * Bc / (t_b * 1e-3) / fp32)
        rows.append({"config": name, "B": Bc, "T": Tc, "nx": s.nx, "nu": s.nu, "solves_per_s": Bc / ((t_f + t_b) * 1e-3), "fwd_ms": t_f,
                     "bwd_ms": t_b, "mean_iters": mi, "max_iters": int(it.max()), "nonpd_frac": float((fw["flags"] & 1).float().mean()),
                     "tight_frac": float(fw["tight"].float().mean()),
                     "fwd_roof_frac": fr_f, "fwd_roof": "hbm" if bf * Bc / hbm >= ff * Bc / fp32 else "fp32",
                     "bwd_roof_frac": fr_b, "bwd_roof": "hbm" if bb * Bc / hbm >= fb * Bc / fp32 else "fp32"})
        del d, fw, w
        torch.cuda.empty_cache()
    return rows


if __name__ == "__main__":
    sys.exit(main())
