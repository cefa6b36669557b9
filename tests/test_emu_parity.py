"""CPU emulation of the CUDA kernels (same per-thread code, tests/emu) against the oracle and the
reference's golden vectors.  Validates the overlap-save / partition / hat-weight / FFT index math
in the GPU-less container; the `-m gpu` tests repeat these through the real kernels."""
import numpy as np
import pytest

from conftest import TOL
from oracle import sonicsim_oracle as so


def bounds_of(idx, P):
    return np.concatenate([[0], np.cumsum(np.bincount(idx, minlength=P - 1))]).astype(np.int32)


def test_forward_spectra_match_numpy(emu):
    rng = np.random.default_rng(0)
    a = rng.standard_normal(4096).astype(np.float32)
    b = np.zeros(4096, np.float32)
    b[:3000] = rng.standard_normal(3000)
    sa, sb = emu.spectra_pair(a, b)
    for src, s in ((a, sa), (b, sb)):
        ref = np.fft.rfft(src.astype(np.float64), 8192) / 8192      # H spectra carry the 1/F of the inverse FFT
        assert abs(s[0].real - ref[0].real) < 1e-6 and abs(s[0].imag - ref[4096].real) < 1e-6   # packed DC / Nyquist
        assert np.abs(s[1:] - ref[1:4096]).max() / np.abs(ref).max() < 2e-6


def test_golden_fixed(emu, golden):
    g = golden("convolve_fixed_receiver")
    for k in range(int(g["n_cases"])):
        y = emu.render(g[f"x{k}"][0], g[f"h{k}"], mode=0)
        assert so.rel_rms(y, g[f"y{k}"]) < TOL


def test_golden_moving_both_trajectory_forms(emu, golden):
    g = golden("convolve_moving_receiver")
    for k in range(int(g["n_cases"])):
        x, h, idx, w = g[f"x{k}"], g[f"h{k}"], g[f"idx{k}"], g[f"w{k}"]
        y_idx = emu.render(x, h, idx=idx, w=w, mode=2)
        y_bnd = emu.render(x, h, bounds=bounds_of(idx, h.shape[0]), mode=1)
        assert so.rel_rms(y_idx, g[f"y{k}"]) < TOL
        # same weights bit for bit (device-side linspace == numpy's); the two forms block the signal
        # differently (waypoint-aligned vs fixed grid), so they agree to fp32 rounding, not bitwise
        assert so.rel_rms(y_bnd, g[f"y{k}"]) < TOL and so.rel_rms(y_idx, y_bnd) < 2e-6


@pytest.mark.parametrize("P,C,L,N", [(3, 1, 1, 100), (2, 2, 4096, 4096), (2, 1, 4097, 8193), (9, 3, 600, 12289),
                                     (5, 2, 9000, 5000), (60, 2, 300, 9000)])
def test_shapes_and_partitions(emu, P, C, L, N):
    rng = np.random.default_rng(P * 1000 + L)
    x, h, pos = so.synth_dry(rng, N), so.synth_rirs(rng, P, C, L), so.synth_path(rng, P)
    np.random.seed(7)
    idx, w = so.setup_dynamic_interp(pos, N)
    ref = so.convolve_moving_exact_f64(x, h, idx, w)
    assert so.rel_rms(emu.render(x, h, idx=idx, w=w, mode=2), ref) < 5e-6
    assert so.rel_rms(emu.render(x, h, bounds=bounds_of(idx, P), mode=1), ref) < 5e-6


def test_zero_length_segments_and_n_lt_segments(emu):
    rng = np.random.default_rng(3)
    P, C, L, N = 9, 2, 40, 5                        # N < S: some segments get no samples
    x, pos = so.synth_dry(rng, N), so.synth_path(rng, P)
    h = rng.standard_normal((P, C, L)).astype(np.float32)      # no leading delay: the first samples are non-trivial
    pos[3] = pos[2]
    for seed in range(50):                      # the reference itself raises for some residue draws
        np.random.seed(seed)
        try:
            idx, w = so.setup_dynamic_interp(pos, N)
            break
        except ValueError:
            continue
    assert len(np.unique(idx)) < P - 1          # some segments are empty
    ref = so.convolve_moving_receiver(x, h, idx, w)
    assert so.rel_rms(emu.render(x, h, bounds=bounds_of(idx, P), mode=1), ref) < TOL


def test_non_monotone_index_arrays(emu):
    """convolve_moving_receiver accepts ANY (idx, w); the indexed mode must honour that."""
    rng = np.random.default_rng(4)
    P, C, L, N = 7, 2, 300, 9000
    x, h = so.synth_dry(rng, N), so.synth_rirs(rng, P, C, L)
    idx = rng.integers(0, P - 1, N)
    w = rng.random(N).astype(np.float32)
    ref = so.convolve_moving_receiver(x, h, idx, w)
    assert so.rel_rms(emu.render(x, h, idx=idx, w=w, mode=2), ref) < TOL


def test_known_answers(emu):
    rng = np.random.default_rng(9)
    N, C, L, P = 10000, 2, 500, 4
    x = so.synth_dry(rng, N)
    # delta RIR -> output = dry
    h = np.zeros((P, C, L), np.float32)
    h[:, :, 0] = 1.0
    np.random.seed(0)
    idx, w = so.setup_dynamic_interp(so.synth_path(rng, P), N)
    y = emu.render(x, h, idx=idx, w=w, mode=2)
    assert so.rel_rms(y, np.stack([x, x])) < 2e-6
    # identical RIRs at both positions -> equals the static convolution
    h1 = so.synth_rirs(rng, 1, C, L)
    h2 = np.repeat(h1, 2, axis=0)
    idx2 = np.zeros(N, np.int64)
    w2 = np.linspace(0, 1, N, endpoint=False).astype(np.float32)
    ym = emu.render(x, h2, idx=idx2, w=w2, mode=2)
    ys = emu.render(x, h1[0], mode=0)
    assert so.rel_rms(ym, ys) < 2e-6


def test_randomised_shapes_around_block_edges(emu):
    """Random shapes biased towards the block / partition edges (4095, 4096, 4097, 8191, ...), both
    trajectory forms and the static path, against the float64 ground truth."""
    rng = np.random.default_rng(2024)
    edges = [1, 2, 255, 256, 257, 4095, 4096, 4097, 8191, 8192, 8193, 12288]
    for case in range(14):
        N = int(rng.choice(edges)) + int(rng.integers(0, 3)) * int(rng.choice([0, 1, 17, 4096]))
        L = int(rng.choice([1, 2, 31, 4095, 4096, 4097, 6000]))
        C = int(rng.integers(1, 4))
        P = int(rng.integers(2, 7))
        N = max(N, P)                                   # at least one sample per ... not required, but keeps counts >= 0
        x = so.synth_dry(rng, N)
        h = rng.standard_normal((P, C, L)).astype(np.float32) * np.exp(-np.arange(L) / max(L / 4, 1)).astype(np.float32)
        pos = so.synth_path(rng, P)
        idx = w = None
        for seed in range(30):
            np.random.seed(seed)
            try:
                idx, w = so.setup_dynamic_interp(pos, N)
                break
            except ValueError:
                continue
        if idx is None:
            continue
        ref = so.convolve_moving_exact_f64(x, h, idx, w)
        tol = 2e-5 if np.sqrt(np.mean(ref ** 2)) > 1e-6 else 1.0
        assert so.rel_rms(emu.render(x, h, idx=idx, w=w, mode=2), ref) < tol, (case, N, L, C, P)
        assert so.rel_rms(emu.render(x, h, bounds=bounds_of(idx, P), mode=1), ref) < tol, (case, N, L, C, P)
        refs = so.convolve_fixed_receiver(x.astype(np.float64)[None], h[0].astype(np.float64))
        assert so.rel_rms(emu.render(x, h[0], mode=0), refs) < tol, (case, N, L, C)
