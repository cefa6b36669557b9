"""The C ABI without Python in the way: tests/cabi/cabi_harness.c is compiled with gcc against include/sonicsim_b200.h,
linked with the in-tree library and run on golden-sized inputs; its output must equal what the ctypes path produces
(same kernels) and match the oracle."""
import os
import shutil
import subprocess

import numpy as np
import pytest

from conftest import TOL
from oracle import sonicsim_oracle as so

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_c_program_against_the_header(tmp_path):
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    from sonicsim_b200 import _lib, render, SonicSim_moving as sm
    exe = str(tmp_path / "cabi_harness")
    csrc = os.path.dirname(_lib.LIB_PATH)
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-O1", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "cabi", "cabi_harness.c"), "-o", exe,
                    "-L", csrc, "-lsonicsim_b200", "-Wl,-rpath," + csrc], check=True)
    rng = np.random.default_rng(31)
    P, C, L, N = 7, 3, 1500, 40000
    x, h, pos = so.synth_dry(rng, N), so.synth_rirs(rng, P, C, L), so.synth_path(rng, P)
    np.random.seed(31)
    idx, w = so.setup_dynamic_interp(pos, N)
    bounds = np.concatenate([[0], np.cumsum(np.bincount(idx, minlength=P - 1))]).astype(np.int32)
    with open(tmp_path / "in.bin", "wb") as f:
        f.write(np.array([N, P, C, L], np.int32).tobytes() + x.tobytes() + h.tobytes() +
                idx.astype(np.int32).tobytes() + w.astype(np.float32).tobytes() + bounds.tobytes())
    res = subprocess.run([exe, str(tmp_path / "in.bin"), str(tmp_path / "out.bin")], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr + res.stdout
    assert "cabi_harness ok" in res.stdout
    out = np.fromfile(tmp_path / "out.bin", dtype=np.float32).reshape(3, C, N)
    ref = so.convolve_moving_receiver(x, h, idx, w)
    assert so.rel_rms(out[0], ref) < TOL
    assert so.rel_rms(out[1], so.convolve_fixed_receiver(x[None], h[0])) < TOL
    assert so.rel_rms(out[2], ref) < TOL
    # same bits as the ctypes route into the same library
    assert np.array_equal(out[0], sm.convolve_moving_receiver(x, h, idx, w))
    assert np.array_equal(out[2], render.default_renderer().render_host([render.MovingSource(x, h, bounds)])[0])
