#!/usr/bin/env python
"""In-tree build of the product's native code (sm_100a only):

  csrc/asg_scan.cu, asg_grad.cu, asg_api.cu  --nvcc-->  libasg_b200.so      (C ABI, include/asg_b200.h)
  csrc/torch_binding.cpp                     --g++--->  torch_asg_native*.so (the reference's module name)

Both land next to this file so they travel with the repo snapshot to the GPU box.  nvcc
cross-compiles without a GPU.  Called by ``__graft_entry__.build()``.
"""
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
INC = os.path.join(HERE, "..", "include")
BUILD = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libasg_b200.so")
CUDA_HOME = os.environ.get("CUDA_HOME", "/usr/local/cuda")
NVCC = os.path.join(CUDA_HOME, "bin", "nvcc")
CXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC", "-ccbin", CXX]
CU_SOURCES = ["asg_scan.cu", "asg_grad.cu", "asg_large.cu", "asg_viterbi.cu", "asg_peer.cu", "asg_replabel.cu", "asg_api.cu"]
HEADERS = [os.path.join(CSRC, "asg_common.cuh"), os.path.join(CSRC, "asg_internal.h"),
           os.path.join(INC, "asg_b200.h")]


def ext_path():
    return os.path.join(HERE, "torch_asg_native" + (sysconfig.get_config_var("EXT_SUFFIX") or ".so"))


def _stale(out, deps):
    return (not os.path.exists(out)) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps)


def build_lib(verbose=False):
    os.makedirs(BUILD, exist_ok=True)
    procs, objs = [], []
    for src in CU_SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(BUILD, src.replace(".cu", ".o"))
        objs.append(o)
        if _stale(o, [s, __file__] + HEADERS):
            cmd = [NVCC] + NVCC_FLAGS + ["-c", s, "-o", o]
            if verbose:
                print(" ".join(cmd), flush=True)
            procs.append((src, subprocess.Popen(cmd)))
    for src, p in procs:
        if p.wait() != 0:
            raise RuntimeError("nvcc failed on " + src)
    if procs or _stale(LIB, objs):
        cmd = [NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs + [
               "-cudart", "shared", "-ccbin", CXX,
               "-Xlinker", "-rpath," + os.path.join(CUDA_HOME, "lib64")]
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.check_call(cmd)
    return LIB


def build_ext(verbose=False):
    import torch
    from torch.utils import cpp_extension as ce
    out = ext_path()
    src = os.path.join(CSRC, "torch_binding.cpp")
    if not _stale(out, [src, LIB, __file__, os.path.join(INC, "asg_b200.h")]):
        return out
    tlib = os.path.join(os.path.dirname(torch.__file__), "lib")
    inc = ce.include_paths() + [sysconfig.get_paths()["include"], os.path.join(CUDA_HOME, "include")]
    cmd = [CXX, "-std=c++17", "-O2", "-fPIC", "-shared", "-w",
           "-DTORCH_EXTENSION_NAME=torch_asg_native", "-DTORCH_API_INCLUDE_EXTENSION_H",
           "-D_GLIBCXX_USE_CXX11_ABI=%d" % int(torch._C._GLIBCXX_USE_CXX11_ABI)]
    cmd += ["-I" + p for p in inc]
    cmd += [src, "-o", out, "-L" + HERE, "-lasg_b200", "-L" + tlib, "-ltorch_python", "-ltorch", "-ltorch_cpu",
            "-lc10", "-lc10_cuda", "-Wl,-rpath,$ORIGIN", "-Wl,-rpath," + tlib,
            "-Wl,-rpath," + os.path.join(CUDA_HOME, "lib64")]
    if verbose:
        print(" ".join(cmd), flush=True)
    subprocess.check_call(cmd)
    return out


def build(verbose=False):
    build_lib(verbose)
    return build_ext(verbose)


if __name__ == "__main__":
    print(build(verbose="-q" not in sys.argv))
