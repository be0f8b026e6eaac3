"""Build libmmer_sm100.so (all CUDA kernels + the C ABI) in-tree with nvcc for sm_100a."""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libmmer_sm100.so")
SOURCES = ["api", "gemm_tc", "gemm_ln", "gemm_simt", "rowops", "attention", "attention_fwd_bf16", "attention_fwd_f32", "attention_bwd_bf16",
           "attention_bwd_f32", "attention_generic", "attention_mma", "attention_long", "ln_pipe", "loss_head", "optim", "bn", "attribution", "batch", "evalops", "serve", "serve_dsmem", "serve_small",
           "engine"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-O3", "-lineinfo",
         "-Xcompiler", "-fPIC"]


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    headers = [os.path.join(CSRC, "common.cuh"), os.path.join(CSRC, "attention_small.cuh"), os.path.join(CSRC, "ptx.cuh"), os.path.join(CSRC, "tc05.cuh"),
               os.path.join(HERE, "..", "include", "mmer.h")]
    jobs = []
    for s in SOURCES:
        src, obj = os.path.join(CSRC, s + ".cu"), os.path.join(OBJ, s + ".o")
        if force or _stale(obj, [src] + headers):
            jobs.append([NVCC] + FLAGS + ["-c", src, "-o", obj])

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed: %s\n%s\n%s" % (" ".join(cmd), r.stdout, r.stderr))
        if verbose and (r.stdout or r.stderr):
            print(r.stdout, r.stderr)

    with ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        list(ex.map(run, jobs))
    objs = [os.path.join(OBJ, s + ".o") for s in SOURCES]
    if force or jobs or _stale(LIB, objs):
        # share the CUDA runtime instance with the host process (PyTorch loads libcudart.so.12 first), so the
        # current device and stream handles mean the same thing on both sides of the C ABI
        rpaths = [os.path.join(p, "nvidia", "cuda_runtime", "lib") for p in sys.path if p.endswith("site-packages")]
        rpaths.append("/usr/local/cuda/lib64")
        link = [NVCC, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a",
                                                       "-cudart", "shared"]
        for r in rpaths:
            link += ["-Xlinker", "-rpath", "-Xlinker", r]
        run(link)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
