"""The tensor-core question at n_x >= 16 (VERDICT r1 missing #4): time the three formulations of Q = C + F'VF in
``csrc/probe_fvf.cu`` with CUDA events and check each against fp64 (developer tool; ncu counters of the same command are
summarised in profiles/r2_probe_fvf.txt).

    python tools/bench_fvf.py [--B 4096] [--reps 50]
"""
import argparse
import ctypes as C
import sys

import torch

sys.path.insert(0, ".")
from tacdmpc_b200 import _lib  # noqa: E402

KINDS = {0: "CTA/element, 1x4 tiles from smem (FFMA2)", 1: "warp/element, lane-per-row registers", 2: "warp/element, mma.sync TF32 x3"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=4096)
    ap.add_argument("--reps", type=int, default=50)
    ap.add_argument("--once", action="store_true", help="one launch per kind (for ncu)")
    a = ap.parse_args()
    dev = torch.device("cuda")
    L = _lib.lib()
    for nx, nu in ((32, 8), (16, 8)):
        n = nx + nu
        g = torch.Generator(device="cpu").manual_seed(nx)
        V = torch.randn(a.B, nx, nx, generator=g)
        V = (V @ V.transpose(1, 2) / nx).contiguous()
        F = (torch.randn(nx, n, generator=g) / nx ** 0.5).contiguous()
        Cm = torch.randn(a.B, n, n, generator=g).contiguous()
        Vd, Fd, Cd = V.to(dev), F.to(dev), Cm.to(dev)
        Q = torch.empty_like(Cd)
        # fp64 answer of one application (reps = 1)
        ref = (Cm.double() + F.double().T @ V.double() @ F.double())
        flop = 2.0 * (nx * nx * n + n * nx * n) * a.B
        for kind, name in KINDS.items():
            def launch(reps):
                _lib.check(L.tacd_probe_fvf(kind, nx, nu, a.B, reps, C.c_void_p(Vd.data_ptr()), C.c_void_p(Fd.data_ptr()),
                                            C.c_void_p(Cd.data_ptr()), C.c_void_p(Q.data_ptr()),
                                            C.c_void_p(torch.cuda.current_stream().cuda_stream)), "tacd_probe_fvf")
            launch(1)
            torch.cuda.synchronize()
            err = ((Q.cpu().double() - ref).abs().max() / ref.abs().max()).item()
            if a.once:
                launch(a.reps)
                torch.cuda.synchronize()
                print(f"nx={nx} nu={nu} kind {kind}: rel err {err:.2e}")
                continue
            ts = []
            for reps in (1, a.reps + 1):
                best = 1e9
                for _ in range(5):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    launch(reps)
                    e1.record()
                    e1.synchronize()
                    best = min(best, e0.elapsed_time(e1))
                ts.append(best)
            per = (ts[1] - ts[0]) / a.reps      # ms per application of the product to the whole batch (staging subtracted)
            print(f"nx={nx} nu={nu} B={a.B} kind {kind} [{name}]: {per * 1e3:8.1f} us per product pass, "
                  f"{flop / per / 1e9:7.2f} TFLOP/s (model flops), rel err vs fp64 {err:.2e}")


if __name__ == "__main__":
    main()
