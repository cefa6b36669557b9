import ctypes
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
TOL = 1e-4          # north-star tolerance: relative RMS vs the reference CPU path


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    def load(name):
        return np.load(os.path.join(GOLDEN, name + ".npz"))
    return load


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def emu():
    """CPU emulation of the CUDA kernels (tests/emu/ss_emu.cu), built with g++."""
    sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
    import build_emu
    lib = ctypes.CDLL(build_emu.build())
    fp, ip = ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_int32)

    def ptr(a, t=fp):
        return a.ctypes.data_as(t) if a is not None else None

    def render(x, h, idx=None, w=None, bounds=None, mode=2):
        x = np.ascontiguousarray(x, np.float32)
        h = np.ascontiguousarray(h, np.float32)
        if h.ndim == 2:
            h = h[None]
        P, C, L = h.shape
        N = x.shape[0]
        out = np.zeros((C, N), np.float32)
        idx = None if idx is None else np.ascontiguousarray(idx, np.int32)
        w = None if w is None else np.ascontiguousarray(w, np.float32)
        bounds = None if bounds is None else np.ascontiguousarray(bounds, np.int32)
        lib.emu_render(ptr(x), ptr(h), ptr(out), ptr(bounds, ip), ptr(idx, ip), ptr(w), N, P, C, L, mode)
        return out

    def spectra_pair(a, b):
        a = np.ascontiguousarray(a, np.float32)
        b = np.ascontiguousarray(b, np.float32)
        sa = np.zeros(8192, np.float32)
        sb = np.zeros(8192, np.float32)
        lib.emu_spectra_pair(ptr(a), ptr(b), a.shape[0], ptr(sa), ptr(sb))
        return sa.view(np.complex64), sb.view(np.complex64)

    def lufs(data, rate, block_size, target=-20.0):
        from sonicsim_b200.SonicSim_audio import gating_plan
        x = np.ascontiguousarray(data, np.float32)
        N = x.shape[0]
        C = 1 if x.ndim == 1 else x.shape[1]
        brk, lo, hi = gating_plan(N, float(rate), float(block_size))
        res = np.zeros(2, np.float64)
        lib.emu_lufs(ptr(x), N, C, ctypes.c_longlong(C), ctypes.c_longlong(1), ctypes.c_double(rate),
                     ctypes.c_double(block_size), ctypes.c_double(target), ptr(brk, ip), len(brk) - 1,
                     ptr(lo, ip), ptr(hi, ip), len(lo), res.ctypes.data_as(ctypes.POINTER(ctypes.c_double)))
        return float(res[0]), float(res[1])

    class Emu:
        pass
    e = Emu()
    e.render = render
    e.spectra_pair = spectra_pair
    e.lufs = lufs
    return e
