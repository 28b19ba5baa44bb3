#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY.  Build recipe for ``oracle/_ref``: the *unmodified* reference
CPU implementation (zh217/torch-asg, ``torch_asg/native/*.cpp``) compiled from the sources
where they lie under ``/root/reference`` into ``oracle/_ref/torch_asg_native_ref*.so``.

Nothing is copied out of the reference tree.  The only compatibility problem with torch 2.x
(the ambiguous ``.sum({0, 1})`` at ``native/fully_connected_lattice.cpp:60``) is handled by
``oracle/ref_wrap/fully_connected_lattice_compat.cpp``, which ``#include``s the reference file
behind a one-line macro.  Flags follow the reference's CPU build (``setup.py:41``:
``-fopenmp -Ofast``).  The module is registered under the name ``torch_asg_native_ref`` so it can
be loaded next to the product's own ``torch_asg_native``.

Two more outputs, same rules (git-ignored, travel with the snapshot, never shipped):

* ``oracle/_ref/ref_cuda/torch_asg_native*.so`` -- the reference's *CUDA* extension (all six sources,
  ``setup.py:13-32`` flags with the arch list swapped for ``sm_100a``), registered under the reference's
  own module name.  ``bench.py`` times it on the B200 as ``gpu_reference`` (the honest same-box GPU
  comparator) in a subprocess whose ``sys.path`` holds only the reference's files.
* ``oracle/_ref/ref_py/`` -- byte-identical copies of the reference's ``torch_asg/__init__.py``,
  ``asg.py`` and ``test/test_asg.py``, staged (not committed: the task forbids reference sources in the
  repo history) so that the ``-m gpu`` suite can run the reference's OWN test file, through the reference's
  OWN ``asg.py``, on top of the product's ``torch_asg_native`` on the GPU box.

Used by: tests/ (to pin the numpy oracle), tools that regenerate tests/golden/, and
``bench.py --impl reference`` / the ``cpu_baseline`` / ``gpu_reference`` legs.  Never by the product path.
"""
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = "/root/reference/torch_asg"
REF_NATIVE = REF_ROOT + "/native"
OUT_DIR = os.path.join(HERE, "_ref")
MODNAME = "torch_asg_native_ref"


def ref_so_path():
    suffix = sysconfig.get_config_var("EXT_SUFFIX") or ".so"
    return os.path.join(OUT_DIR, MODNAME + suffix)


def build(force=False, verbose=False):
    """Compile the reference CPU extension.  Returns the .so path, or None when the reference
    tree is not present (GPU box: the prebuilt file travels with the snapshot)."""
    so = ref_so_path()
    if not os.path.isdir(REF_NATIVE):
        return so if os.path.exists(so) else None
    srcs = [
        os.path.join(REF_NATIVE, "utils.cpp"),
        os.path.join(REF_NATIVE, "force_aligned_lattice.cpp"),
        os.path.join(HERE, "ref_wrap", "fully_connected_lattice_compat.cpp"),
        os.path.join(REF_NATIVE, "extension.cpp"),
    ]
    if (not force) and os.path.exists(so) and all(
            os.path.getmtime(so) >= os.path.getmtime(s) for s in srcs + [__file__]):
        return so
    os.makedirs(OUT_DIR, exist_ok=True)
    import torch
    from torch.utils import cpp_extension as ce
    inc = ce.include_paths() + [sysconfig.get_paths()["include"], REF_NATIVE]
    libdir = os.path.join(os.path.dirname(torch.__file__), "lib")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    common = [cxx, "-std=c++17", "-fPIC", "-fopenmp", "-Ofast", "-w",
              "-DTORCH_EXTENSION_NAME=" + MODNAME, "-DTORCH_API_INCLUDE_EXTENSION_H",
              "-D_GLIBCXX_USE_CXX11_ABI=%d" % int(torch._C._GLIBCXX_USE_CXX11_ABI)]
    common += ["-I" + p for p in inc]
    objs = []
    procs = []
    for s in srcs:
        o = os.path.join(OUT_DIR, os.path.basename(s).replace(".cpp", ".o"))
        objs.append(o)
        cmd = common + ["-c", s, "-o", o]
        if verbose:
            print(" ".join(cmd))
        procs.append(subprocess.Popen(cmd))
    for p in procs:
        if p.wait() != 0:
            raise RuntimeError("reference compile failed")
    link = [cxx, "-shared", "-fopenmp"] + objs + ["-o", so, "-L" + libdir,
            "-Wl,-rpath," + libdir, "-ltorch_python", "-ltorch", "-ltorch_cpu", "-lc10"]
    if verbose:
        print(" ".join(link))
    subprocess.check_call(link)
    for o in objs:
        os.remove(o)
    return so


def ref_cuda_dir():
    return os.path.join(OUT_DIR, "ref_cuda")


def ref_cuda_so_path():
    suffix = sysconfig.get_config_var("EXT_SUFFIX") or ".so"
    return os.path.join(ref_cuda_dir(), "torch_asg_native" + suffix)


def build_cuda(force=False, verbose=False):
    """Compile the reference's CUDA extension (module name ``torch_asg_native``, as its asg.py imports it)
    for sm_100a.  Returns the .so path, or None when the reference tree is not present and nothing
    prebuilt travelled."""
    so = ref_cuda_so_path()
    if not os.path.isdir(REF_NATIVE):
        return so if os.path.exists(so) else None
    cpp = [
        os.path.join(REF_NATIVE, "utils.cpp"),
        os.path.join(REF_NATIVE, "force_aligned_lattice.cpp"),
        os.path.join(HERE, "ref_wrap", "fully_connected_lattice_compat.cpp"),
        os.path.join(REF_NATIVE, "extension.cpp"),
        os.path.join(HERE, "ref_wrap", "streamlined_fast_gpu_compat.cpp"),
    ]
    cu = os.path.join(REF_NATIVE, "force_aligned_lattice_kernel.cu")
    if (not force) and os.path.exists(so) and all(
            os.path.getmtime(so) >= os.path.getmtime(s) for s in cpp + [cu, __file__]):
        return so
    os.makedirs(ref_cuda_dir(), exist_ok=True)
    import torch
    from torch.utils import cpp_extension as ce
    cuda_home = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    inc = ce.include_paths() + [sysconfig.get_paths()["include"], REF_NATIVE, os.path.join(cuda_home, "include")]
    libdir = os.path.join(os.path.dirname(torch.__file__), "lib")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    defs = ["-DTORCH_EXTENSION_NAME=torch_asg_native", "-DTORCH_API_INCLUDE_EXTENSION_H", "-DTORCH_ASG_SUPPORTS_CUDA",
            "-D_GLIBCXX_USE_CXX11_ABI=%d" % int(torch._C._GLIBCXX_USE_CXX11_ABI)]
    incs = ["-I" + p for p in inc]
    procs, objs = [], []
    for s in cpp:      # setup.py:20-22
        o = os.path.join(ref_cuda_dir(), os.path.basename(s).replace(".cpp", ".o"))
        objs.append(o)
        cmd = [cxx, "-std=c++17", "-fPIC", "-O2", "-fopenmp", "-w"] + defs + incs + ["-c", s, "-o", o]
        if verbose:
            print(" ".join(cmd))
        procs.append(subprocess.Popen(cmd))
    o = os.path.join(ref_cuda_dir(), "force_aligned_lattice_kernel.o")
    objs.append(o)
    # setup.py:23-32 with the sm_60..sm_75 list replaced by the one arch this box has
    cmd = [os.path.join(cuda_home, "bin", "nvcc"), "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "--use_fast_math",
           "--expt-extended-lambda", "--expt-relaxed-constexpr", "-std=c++17", "-w", "-Xcompiler", "-fPIC", "-ccbin", cxx] + \
        defs + incs + ["-c", cu, "-o", o]
    if verbose:
        print(" ".join(cmd))
    procs.append(subprocess.Popen(cmd))
    for p in procs:
        if p.wait() != 0:
            raise RuntimeError("reference CUDA compile failed")
    link = [cxx, "-shared", "-fopenmp"] + objs + ["-o", so, "-L" + libdir, "-L" + os.path.join(cuda_home, "lib64"),
            "-Wl,-rpath," + libdir, "-Wl,-rpath," + os.path.join(cuda_home, "lib64"),
            "-ltorch_python", "-ltorch", "-ltorch_cpu", "-ltorch_cuda", "-lc10", "-lc10_cuda", "-lcudart"]
    if verbose:
        print(" ".join(link))
    subprocess.check_call(link)
    for o in objs:
        os.remove(o)
    return so


def ref_py_dir():
    return os.path.join(OUT_DIR, "ref_py")


def stage_ref_py():
    """Byte-identical copies of the reference's Python surface and its test file under oracle/_ref/ref_py
    (git-ignored).  Returns the directory, or None when neither the reference nor a staged copy exists."""
    import filecmp
    import shutil
    dst = ref_py_dir()
    files = [("__init__.py", "torch_asg/__init__.py"), ("asg.py", "torch_asg/asg.py"),
             ("test/test_asg.py", "test_asg.py")]
    if not os.path.isdir(REF_ROOT):
        return dst if all(os.path.exists(os.path.join(dst, d)) for _, d in files) else None
    os.makedirs(os.path.join(dst, "torch_asg"), exist_ok=True)
    for src, d in files:
        a, b = os.path.join(REF_ROOT, src), os.path.join(dst, d)
        if not (os.path.exists(b) and filecmp.cmp(a, b, shallow=False)):
            shutil.copyfile(a, b)
    return dst


def load():
    """Import and return the compiled reference module (torch must be imported first)."""
    import importlib.util
    import torch  # noqa: F401  (libtorch symbols)
    so = build()
    if so is None or not os.path.exists(so):
        raise ImportError("oracle/_ref is not built and /root/reference is not available")
    spec = importlib.util.spec_from_file_location(MODNAME, so)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose=True)
    print("built" if p else "reference tree not available", p)
    print("staged", stage_ref_py())
    if "--cuda" in sys.argv:
        print("cuda", build_cuda(force="--force" in sys.argv, verbose=True))
