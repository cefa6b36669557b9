"""The oracle (oracle/sonicsim_oracle.py) against the golden vectors produced by the UNMODIFIED
reference (oracle/make_golden.py), and against the live reference when /root/reference exists."""
import numpy as np
import pytest

from oracle import ref_loader
from oracle import sonicsim_oracle as so


def test_setup_dynamic_interp_bit_exact(golden):
    g = golden("setup_dynamic_interp")
    for k in range(int(g["n_cases"])):
        np.random.seed(int(g[f"seed{k}"]))
        idx, w = so.setup_dynamic_interp(g[f"pos{k}"], int(g[f"N{k}"]))
        assert idx.dtype == np.int64 and w.dtype == np.float32
        assert np.array_equal(idx, g[f"idx{k}"])
        assert np.array_equal(w, g[f"w{k}"])


def test_convolve_fixed_receiver(golden):
    g = golden("convolve_fixed_receiver")
    for k in range(int(g["n_cases"])):
        y = so.convolve_fixed_receiver(g[f"x{k}"], g[f"h{k}"])
        assert y.shape == g[f"y{k}"].shape
        assert so.rel_rms(y, g[f"y{k}"]) < 1e-6


def test_convolve_moving_receiver(golden):
    g = golden("convolve_moving_receiver")
    for k in range(int(g["n_cases"])):
        y = so.convolve_moving_receiver(g[f"x{k}"], g[f"h{k}"], g[f"idx{k}"].astype(np.int64), g[f"w{k}"])
        assert y.dtype == np.float32
        assert so.rel_rms(y, g[f"y{k}"]) < 1e-6
        # the segment-local float64 form is the same function
        y64 = so.convolve_moving_exact_f64(g[f"x{k}"], g[f"h{k}"], g[f"idx{k}"], g[f"w{k}"])
        assert so.rel_rms(g[f"y{k}"], y64) < 5e-6


def test_interpolate_moving_audio(golden):
    g = golden("interpolate_moving_audio")
    for k in range(int(g["n_cases"])):
        np.random.seed(int(g[f"seed{k}"]))
        y = so.interpolate_moving_audio(g[f"x{k}"], g[f"h{k}"], g[f"pos{k}"])
        assert so.rel_rms(y, g[f"y{k}"]) < 1e-6


def test_fft_conv(golden):
    import torch
    g = golden("fft_conv")
    for k in range(int(g["n_cases"])):
        y = so.fft_conv(torch.from_numpy(g[f"x{k}"]), torch.from_numpy(g[f"h{k}"])).numpy()
        assert y.shape == g[f"y{k}"].shape
        assert so.rel_rms(y, g[f"y{k}"]) < 1e-6


def test_degenerate_paths_raise_like_reference():
    # all-identical waypoints -> divide by zero -> ValueError from np.repeat
    pos = np.ones((4, 3))
    with np.errstate(all="ignore"):
        with pytest.raises(ValueError):
            so.setup_dynamic_interp(pos, 100)


@pytest.mark.skipif(not ref_loader.available(), reason="/root/reference not present (GPU box)")
def test_oracle_equals_live_reference():
    ref = ref_loader.load()
    rng = np.random.default_rng(5)
    x = so.synth_dry(rng, 12000)
    h = so.synth_rirs(rng, 6, 2, 900)
    pos = so.synth_path(rng, 6)
    np.random.seed(3)
    i1, w1 = ref.setup_dynamic_interp(pos, 12000)
    np.random.seed(3)
    i2, w2 = so.setup_dynamic_interp(pos, 12000)
    assert np.array_equal(i1, i2) and np.array_equal(w1, w2)
    assert np.array_equal(ref.convolve_moving_receiver(x, h, i1, w1), so.convolve_moving_receiver(x, h, i2, w2))
    assert np.array_equal(ref.convolve_fixed_receiver(x[None], h[0]), so.convolve_fixed_receiver(x[None], h[0]))
