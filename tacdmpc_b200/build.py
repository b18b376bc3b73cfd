"""Build the CUDA library (tacdmpc_b200/libtacd_ilqr.so) in-tree for sm_100a.

    python -m tacdmpc_b200.build [--force]

nvcc cross-compiles without a GPU.  The .so is git-ignored but travels to the GPU box
with the gpurun snapshot; there is no JIT at import time.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libtacd_ilqr.so")
SOURCES = ["capi.cu", "ilqr_small_f32.cu", "ilqr_small_f64.cu", "ilqr_group_f32.cu", "ilqr_group_f64.cu", "ilqr_generic.cu", "ilqr_generic_part1.cu", "ilqr_generic_part2.cu", "ilqr_generic_part3.cu",
           "ppo_glue.cu", "probe.cu", "probe_fvf.cu", "ilqr_lti_f32.cu", "ilqr_lti_f64.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]


def _nvcc() -> str:
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")


def _deps():
    out = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    out.append(os.path.join(os.path.dirname(HERE), "include", "tacd_ilqr.h"))
    return out


def is_stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in _deps())


def build_extension(force: bool = False, verbose: bool = False, defines=(), out: str | None = None, sources=None) -> str:
    """Compile the library.  `defines` / `out` build an experiment variant (e.g. ("-DTACD_GROUP_MAXW=16",) into
    tacdmpc_b200/variants/x.so, selected at run time with TACD_LIB=path); the default call builds the product."""
    variant = bool(defines) or out is not None
    if not variant and not force and not is_stale():
        return LIB
    nvcc = _nvcc()
    objdir = os.path.join(HERE, "build" if not variant else "build_" + os.path.splitext(os.path.basename(out or "variant"))[0])
    os.makedirs(objdir, exist_ok=True)
    env = dict(os.environ)
    # the image exports CC=/opt/gcc/bin/gcc; nvcc wants the system g++
    ccbin = ["-ccbin", "/usr/bin/g++"] if os.path.exists("/usr/bin/g++") else []

    def compile_one(src):
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, *defines, *ccbin, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True, env=env)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    target = out or LIB
    os.makedirs(os.path.dirname(target), exist_ok=True)
    cmd = [nvcc, "-shared", *ccbin, "-gencode", "arch=compute_100a,code=sm_100a", "-o", target, *objs]
    r = subprocess.run(cmd, capture_output=True, text=True, env=env)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return target


if __name__ == "__main__":
    defs = tuple(a for a in sys.argv[1:] if a.startswith("-D"))
    outs = [a.split("=", 1)[1] for a in sys.argv[1:] if a.startswith("--out=")]
    print(build_extension(force="--force" in sys.argv, verbose="-v" in sys.argv, defines=defs, out=outs[0] if outs else None))
