"""CPU tier: the C-ABI shared library loads without a GPU and exports every symbol that
include/asg_b200.h declares; argument validation works; compute entry points refuse to run without a
device (there is no CPU fallback).  No kernels are launched here."""
import ctypes
import os
import re

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
HEADER = os.path.join(ROOT, "include", "asg_b200.h")
LIB = os.path.join(ROOT, "torch-asg_b200", "libasg_b200.so")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(asg_[a-z_0-9]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(LIB):
        import __graft_entry__ as g
        g.build()
    lib = ctypes.CDLL(LIB)
    lib.asg_last_error.restype = ctypes.c_char_p
    lib.asg_version.restype = ctypes.c_char_p
    lib.asg_kernel_name.restype = ctypes.c_char_p
    return lib


def test_every_declared_symbol_is_exported(lib):
    names = declared_symbols()
    assert {"asg_layout", "asg_scan", "asg_grad", "asg_host_create", "asg_host_destroy", "asg_host_loss", "asg_host_submit", "asg_host_wait",
            "asg_step", "asg_forward", "asg_backward", "asg_release", "asg_last_error", "asg_version", "asg_launch_count", "asg_profile_enable", "asg_profile_select", "asg_profile_stride", "asg_profile_read"} <= set(names)
    for n in names:
        assert hasattr(lib, n), "libasg_b200.so does not export " + n
    assert b"sm_100a" in lib.asg_version()


def test_sass_is_sm_100a_only():
    import subprocess
    out = subprocess.run(["/usr/local/cuda/bin/cuobjdump", "-lelf", LIB], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


class Layout(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int64) for n in ("sp", "nblk", "nbp", "lpi", "tp")] + [
        (n, ctypes.c_size_t) for n in ("elem_bytes", "rec_bytes", "state_bytes", "ctx_bytes", "grad_scratch_bytes",
                                       "rec_fac_alpha", "rec_fac_beta", "rec_fac_ea", "rec_fac_eb", "rec_fcc_alpha",
                                       "rec_fcc_beta", "rec_fcc_w", "rec_fcc_scale",
                                       "off_ehat", "off_fac_hd", "off_lengths", "off_scalars", "off_ehat_t", "off_fmax", "off_rmax", "off_large_ws")]


def _layout(lib, dt, T, N, L, S):
    lay = Layout()
    rc = lib.asg_layout(dt, ctypes.c_int64(T), ctypes.c_int64(N), ctypes.c_int64(L), ctypes.c_int64(S), ctypes.byref(lay))
    return rc, lay


def test_layout_and_validation(lib):
    rc, lay = _layout(lib, 0, 1000, 64, 40, 200)
    assert rc == 0
    assert (lay.sp, lay.nblk, lay.nbp, lay.lpi, lay.tp, lay.elem_bytes) == (200, 25, 32, 40, 1000, 4)
    parts = [lay.rec_fac_alpha, lay.rec_fac_beta, lay.rec_fac_ea, lay.rec_fac_eb, lay.rec_fcc_alpha, lay.rec_fcc_beta,
             lay.rec_fcc_w, lay.rec_fcc_scale]
    # one frame record: FAC alpha/beta (2 x 800 B), their exponents (2 x 128 B), FCC alpha/beta/u*beta (3 x 160 B), scale
    assert parts == [0, 800, 1600, 1728, 1856, 2016, 2176, 2336] and lay.rec_bytes == 2368
    assert all(p % 32 == 0 for p in parts) and lay.rec_bytes % 32 == 0
    assert lay.state_bytes == 64 * 1000 * 2368
    offs = [lay.off_ehat, lay.off_fac_hd, lay.off_lengths, lay.off_scalars]
    assert offs == sorted(offs) and all(o % 256 == 0 for o in offs) and lay.ctx_bytes > offs[-1]
    rc, lay = _layout(lib, 1, 7, 3, 5, 3)
    assert rc == 0 and (lay.sp, lay.nblk, lay.nbp, lay.lpi, lay.tp, lay.elem_bytes) == (8, 1, 8, 8, 8, 8)
    assert lay.rec_bytes == 64 + 64 + 32 + 32 + 3 * 64 + 32 and lay.state_bytes == 3 * 8 * lay.rec_bytes
    assert _layout(lib, 5, 4, 1, 3, 2)[0] == -1 and b"dtype" in lib.asg_last_error()
    assert _layout(lib, 0, 4, 1, 3, 5)[0] == -2          # S > T
    assert _layout(lib, 0, 4, 0, 3, 2)[0] == -2
    assert lib.asg_layout(0, ctypes.c_int64(4), ctypes.c_int64(1), ctypes.c_int64(3), ctypes.c_int64(2), None) == -3


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("this check is for the GPU-less box")
    h = ctypes.c_void_p()
    rc = lib.asg_host_create(0, 0, ctypes.c_int64(8), ctypes.c_int64(1), ctypes.c_int64(4), ctypes.c_int64(2), ctypes.byref(h))
    assert rc == -5 and b"no CPU fallback" in lib.asg_last_error()
    import torch_asg_native as nat
    x = torch.zeros(4, 2, 3)
    with pytest.raises(RuntimeError, match="no CPU implementation"):
        nat.fully_connected_forward(x, torch.zeros(3, 3), torch.full((2,), 4), 4, 2, 3)
    from torch_asg import ASGLoss
    with pytest.raises(RuntimeError, match="no CPU implementation"):
        ASGLoss(3)(x, torch.zeros(2, 2, dtype=torch.long))


def test_kernel_names(lib):
    n = lib.asg_kernel_count()
    names = [lib.asg_kernel_name(i).decode() for i in range(n)]
    assert "grad_kernel" in names and "fcc_scan_kernel" in names and "fac_scan_kernel" in names
