"""Builds libfdts_b200.so in-tree with nvcc for sm_100a (no JIT cache: the .so
travels to the GPU box with the repository snapshot)."""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
SOURCES = ["fdts_api.cu", "pattern.cu", "inst_heat.cu", "inst_bbm.cu", "inst_burgers.cu", "inst_ch.cu", "inst_diffusion.cu",
           "kernels_heat.cu", "kernels_heat2.cu", "microbench.cu", "linalg.cu", "gather_plan.cu", "kernels_dmma.cu", "halo_nccl.cu"]
HEADERS = ["fdts_common.cuh", "kernels_generic.cuh", "kargs.cuh", "ref_tables.cuh", "heat_common.cuh",
           "../../include/fdts.h"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
         "--expt-relaxed-constexpr", "-Xptxas", "-v"]
LIB = os.path.join(HERE, "libfdts_b200.so")


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _compile(src):
    obj = os.path.join(HERE, "build", src.replace(".cu", ".o"))
    deps = [os.path.join(HERE, src)] + [os.path.join(HERE, h) for h in HEADERS]
    if _stale(obj, deps):
        log = obj + ".log"
        with open(log, "w") as f:
            subprocess.check_call([NVCC, *FLAGS, "-c", os.path.join(HERE, src), "-o", obj], stdout=f, stderr=f)
    return obj


def build_dev(verbose: bool = False) -> str:
    """development build with the timing switches compiled in (-DFDTS_DEV_SWITCHES): libfdts_b200_dev.so,
    used only by tools/ through FDTS_B200_LIB; its results are wrong by design when a switch is on."""
    # FDTS_DEV_DEFS="-DX=1 -DY=2" FDTS_DEV_TAG=x: an experiment variant (own object directory and library name)
    extra = os.environ.get("FDTS_DEV_DEFS", "").split()
    tag = os.environ.get("FDTS_DEV_TAG", "")
    bdir = os.path.join(HERE, "build", "dev" + tag)
    os.makedirs(bdir, exist_ok=True)
    lib = os.path.join(HERE, f"libfdts_b200_dev{tag}.so")

    only = os.environ.get("FDTS_DEV_SRCS", "").split()  # variant: recompile these, take the rest from build/dev

    def comp(src):
        if tag and only and src not in only:
            return os.path.join(HERE, "build", "dev", src.replace(".cu", ".o"))
        obj = os.path.join(bdir, src.replace(".cu", ".o"))
        deps = [os.path.join(HERE, src)] + [os.path.join(HERE, h) for h in HEADERS]
        if _stale(obj, deps):
            with open(obj + ".log", "w") as f:
                subprocess.check_call([NVCC, *FLAGS, "-DFDTS_DEV_SWITCHES", *extra, "-c", os.path.join(HERE, src), "-o", obj],
                                      stdout=f, stderr=f)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(comp, SOURCES))
    if _stale(lib, objs):
        subprocess.check_call([NVCC, "-shared", "-o", lib, *objs, "-gencode", "arch=compute_100a,code=sm_100a",
                               "-lcudart", "-ldl"])
    return lib


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    if force:
        for f in os.listdir(os.path.join(HERE, "build")):
            if os.path.isfile(os.path.join(HERE, "build", f)):
                os.remove(os.path.join(HERE, "build", f))
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(_compile, SOURCES))
    if _stale(LIB, objs):
        subprocess.check_call([NVCC, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a",
                               "-lcudart", "-ldl"])
        if verbose:
            print("linked", LIB)
    return LIB


if __name__ == "__main__":
    if "--dev" in sys.argv:
        print(build_dev(verbose=True))
    else:
        build(force="--force" in sys.argv, verbose=True)
