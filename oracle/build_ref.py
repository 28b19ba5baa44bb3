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

Used by: tests/ (to pin the numpy oracle), tools that regenerate tests/golden/, and
``bench.py --impl reference`` / the ``cpu_baseline`` leg.  Never by the product path.
"""
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
REF_NATIVE = "/root/reference/torch_asg/native"
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
