"""Build librcdm_b200.so in-tree with nvcc for sm_100a (no torch involved).

    python -m rcdms_b200.build [--force]

Sources: rcdms_b200/csrc/*.cu  ->  rcdms_b200/_C/librcdm_b200.so  (git-ignored; ships to the GPU box with the
snapshot).  cudart is linked statically and the driver API entry point for TMA descriptors is resolved at
run time, so the library loads (and exports every symbol of include/rcdm.h) on a machine without a GPU.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT_DIR = os.environ.get("RCDM_BUILD_DIR") or os.path.join(HERE, "_C")  # RCDM_BUILD_DIR: experiment variants only
LIB = os.path.join(OUT_DIR, "librcdm_b200.so")
SOURCES = ["gemm_host.cu", "attention_host.cu", "norm_host.cu", "unet.cu", "api.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--use_fast_math",
              "-diag-suppress", "177"]
# IEEE arithmetic (no --use_fast_math) for the translation units that hold the scheduler-step kernels, whose results
# are specified bit for bit against torch's op sequence (ddim_cfg_step_kernel, unclip_cfg_step_kernel), and the
# timestep sinusoid (sinf / cosf of arguments up to ~1000 rad).  Their hot loops use explicit intrinsics where wanted.
IEEE_SOURCES = {"api.cu", "unet.cu"}


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def _digest() -> str:
    h = hashlib.sha256()
    for root in (CSRC, os.path.join(os.path.dirname(HERE), "include")):
        for f in sorted(os.listdir(root)):
            if f.endswith((".cu", ".cuh", ".h")):
                h.update(f.encode())
                h.update(open(os.path.join(root, f), "rb").read())
    h.update(" ".join(NVCC_FLAGS + sorted(IEEE_SOURCES)).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = True) -> str:
    global NVCC_FLAGS
    extra = os.environ.get("RCDM_EXTRA_NVCC_FLAGS", "").split()
    if extra:
        NVCC_FLAGS = [f for f in NVCC_FLAGS if f not in extra] + extra
    os.makedirs(OUT_DIR, exist_ok=True)
    stamp = os.path.join(OUT_DIR, "build.sha256")
    dig = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read().strip() == dig:
        return LIB
    nvcc = _nvcc()

    def compile_one(src: str) -> str:
        obj = os.path.join(OUT_DIR, src.replace(".cu", ".o"))
        flags = [f for f in NVCC_FLAGS if not (src in IEEE_SOURCES and f == "--use_fast_math")]
        cmd = [nvcc, *flags, "-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        return obj

    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 4)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    cmd = [nvcc, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static",
           "-lpthread", "-ldl", "-lrt"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    with open(stamp, "w") as f:
        f.write(dig)
    if verbose:
        print(f"built {LIB} ({os.path.getsize(LIB) / 1e6:.1f} MB)")
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv)
