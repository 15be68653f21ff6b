"""Builds the in-tree native libraries for sm_100a.

  libgai_b200.so   hand-written CUDA kernels + the C ABI declared in include/gai_b200.h   (nvcc)
  libgai_host.so   C++ host mirror of the reference's layer/model API over that ABI       (g++)

Everything is compiled in-tree so the built objects travel with the repository snapshot to the GPU box.
"""
from __future__ import annotations

import glob
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "graphaibench_b200")
CSRC = os.path.join(PKG, "csrc")
HOST = os.path.join(PKG, "host")
LIB = os.path.join(PKG, "libgai_b200.so")
HOSTLIB = os.path.join(PKG, "libgai_host.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build_cuda(force=False, verbose=False):
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    hdrs = glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(ROOT, "include", "*.h"))
    objdir = os.path.join(PKG, "build")
    os.makedirs(objdir, exist_ok=True)
    objs, procs = [], []
    for s in srcs:
        o = os.path.join(objdir, os.path.basename(s)[:-3] + ".o")
        objs.append(o)
        if force or _newer(o, [s] + hdrs):
            cmd = [NVCC, *ARCH, "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "-I", os.path.join(ROOT, "include"),
                   "-I", CSRC, "-c", s, "-o", o]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
            for d in os.environ.get("GAI_NVCC_DEFINES", "").split():  # build-time experiments, e.g. GAI_NVCC_DEFINES=-DGAI_GATHER_NOALLOC
                cmd.insert(1, d)
            procs.append((s, subprocess.Popen(cmd)))
    for s, p in procs:
        if p.wait() != 0:
            raise RuntimeError(f"nvcc failed on {s}")
    if force or procs or _newer(LIB, objs):
        subprocess.check_call([NVCC, *ARCH, "-shared", "-o", LIB, *objs, "-lcudart"])
    return LIB


def build_host(force=False):
    srcs = sorted(glob.glob(os.path.join(HOST, "*.cpp")))
    if not srcs:
        return None
    hdrs = glob.glob(os.path.join(HOST, "*.h")) + glob.glob(os.path.join(ROOT, "include", "*.h"))
    if force or _newer(HOSTLIB, srcs + hdrs + [LIB]):
        lib_srcs = [s for s in srcs if not os.path.basename(s).startswith(("train", "convert_main"))]
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-I", os.path.join(ROOT, "include"), "-I", HOST,
                               *lib_srcs, "-o", HOSTLIB, "-L", PKG, "-lgai_b200", "-Wl,-rpath,$ORIGIN"])
        for arch, macro in (("gcn", []), ("sage", ["-DUSE_SAGE"]), ("gat", ["-DUSE_GAT"])):
            subprocess.check_call(["g++", "-O2", "-std=c++17", *macro, "-I", os.path.join(ROOT, "include"), "-I", HOST,
                                   os.path.join(HOST, "train.cpp"), "-o", os.path.join(PKG, f"gpu_train_{arch}"),
                                   "-L", PKG, "-lgai_host", "-lgai_b200", "-Wl,-rpath,$ORIGIN"])
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-I", os.path.join(ROOT, "include"), "-I", HOST, os.path.join(HOST, "convert_main.cpp"),
                               "-o", os.path.join(PKG, "gpu_converter"), "-L", PKG, "-lgai_host", "-lgai_b200", "-Wl,-rpath,$ORIGIN"])
    return HOSTLIB


def build_all(force=False, verbose=False):
    build_cuda(force, verbose)
    build_host(force)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print("built", LIB)
