"""Host-side logic of the package (no GPU): the trajectory setup mirror, argument handling, and
that the C-ABI library loads and exports every symbol include/sonicsim_b200.h declares."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT
from sonicsim_b200 import SonicSim_moving as sm
from sonicsim_b200 import _lib


def test_setup_dynamic_interp_bit_exact_vs_reference_golden(golden):
    g = golden("setup_dynamic_interp")
    for k in range(int(g["n_cases"])):
        np.random.seed(int(g[f"seed{k}"]))
        idx, w = sm.setup_dynamic_interp(g[f"pos{k}"], int(g[f"N{k}"]))
        assert idx.dtype == np.int64 and w.dtype == np.float32
        assert np.array_equal(idx, g[f"idx{k}"]) and np.array_equal(w, g[f"w{k}"])


def test_bounds_are_the_compact_form(golden):
    g = golden("setup_dynamic_interp")
    for k in range(int(g["n_cases"])):
        np.random.seed(int(g[f"seed{k}"]))
        spi = sm._samples_per_interval(g[f"pos{k}"], int(g[f"N{k}"]))
        b = sm.bounds_from_counts(spi)
        assert b.dtype == np.int32 and b[0] == 0 and b[-1] == int(g[f"N{k}"]) and len(b) == len(g[f"pos{k}"])
        assert np.array_equal(np.repeat(np.arange(len(spi)), spi), g[f"idx{k}"])


def test_degenerate_path_raises_value_error():
    with np.errstate(all="ignore"):
        with pytest.raises(ValueError):
            sm.setup_dynamic_interp(np.ones((4, 3)), 100)


def test_library_exports_every_declared_symbol():
    _lib.build()
    hdr = open(os.path.join(ROOT, "include", "sonicsim_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = sorted(set(re.findall(r"\b(ss_[a-z_0-9]+)\s*\(", hdr)))
    assert declared, "no declarations parsed"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), "library does not export %s" % name
    assert sorted(_lib.EXPORTS) == declared
    assert lib.ss_version() >= 100


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError):
        sm.convolve_fixed_receiver(np.zeros((1, 100), np.float32), np.zeros((1, 10), np.float32))


def test_index_error_like_reference():
    x = np.zeros(10, np.float32)
    h = np.zeros((3, 1, 4), np.float32)
    with pytest.raises(IndexError):
        sm.convolve_moving_receiver(x, h, np.full(10, 2), np.zeros(10, np.float32))   # idx + 1 == P


def test_ctypes_structs_match_the_c_header_layout(tmp_path):
    """The ctypes mirrors in _lib.py against include/sonicsim_b200.h as gcc lays it out: same size and the same
    offset for every field (the header is the contract a maintainer binds against)."""
    import ctypes
    import shutil
    import subprocess
    from sonicsim_b200 import _lib
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    pairs = {"ss_source": _lib.SsSource, "ss_loud_item": _lib.SsLoudItem, "ss_post_lufs": _lib.SsPostLufs,
             "ss_mix_item": _lib.SsMixItem}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "sonicsim_b200.h"', 'int main(void) {']
    for cname, cls in pairs.items():
        lines.append('printf("%s size %%zu\\n", sizeof(%s));' % (cname, cname))
        for fname, _ in cls._fields_:
            lines.append('printf("%s %s %%zu\\n", offsetof(%s, %s));' % (cname, fname, cname, fname))
    lines += ['return 0;', '}']
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    inc = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", inc, str(src), "-o", str(exe)], check=True)     # the header is plain C
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split("\n")
    got = {tuple(l.split()[:2]): int(l.split()[2]) for l in out if l.strip()}
    for cname, cls in pairs.items():
        assert got[(cname, "size")] == ctypes.sizeof(cls), cname
        for fname, _ in cls._fields_:
            assert got[(cname, fname)] == getattr(cls, fname).offset, (cname, fname)
