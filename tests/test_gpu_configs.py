"""BASELINE.json configs at their full sizes, through size-independent properties where the oracle
would take minutes: linearity in the RIR set, delta-RIR identity, static == moving with identical
RIRs, batch == single calls.  (configs[1] is checked against the oracle directly in
test_gpu_parity.py; configs[0] below.)"""
import numpy as np
import pytest
import torch

from conftest import TOL
from oracle import sonicsim_oracle as so

pytestmark = pytest.mark.gpu


def test_cfg1_static_mono_4s_against_reference_path():
    """configs[0]: 1 static source, 1 mono mic, 4096-tap RIR, 4 s @ 16 kHz (scipy.fftconvolve reference)."""
    from sonicsim_b200 import SonicSim_moving as sm
    rng = np.random.default_rng(1000)
    x, h = so.synth_dry(rng, 64000), so.synth_rirs(rng, 1, 1, 4096)[0]
    y = sm.convolve_fixed_receiver(x[None], h)
    assert y.shape == (1, 64000)
    assert so.rel_rms(y, so.convolve_fixed_receiver(x[None], h)) < TOL


def test_cfg3_scene_batch_60s_60pt():
    """configs[2] (one utterance of the 64): 2 moving (P=60) + 2 static sources, 6 mics, 60 s; checked
    against the oracle for one moving and one static stem, the rest via batch == single-call identity."""
    from sonicsim_b200 import SonicSim_moving as sm
    from sonicsim_b200 import render
    rng = np.random.default_rng(3000)
    N, C, L, P = 960000, 6, 4096, 60
    moving = [(so.synth_dry(rng, N), so.synth_rirs(rng, P, C, L), so.synth_path(rng, P)) for _ in range(2)]
    static = [(so.synth_dry(rng, N), so.synth_rirs(rng, 1, C, L)[0]) for _ in range(2)]
    np.random.seed(3000)
    ym, ys = render.render_scene(moving, static)
    np.random.seed(3000)
    idx, w = so.setup_dynamic_interp(moving[0][2], N)
    assert so.rel_rms(ym[0], so.convolve_moving_receiver(moving[0][0], moving[0][1], idx, w)) < TOL
    assert so.rel_rms(ys[1], so.convolve_fixed_receiver(static[1][0][None], static[1][1])) < TOL
    idx2, w2 = so.setup_dynamic_interp(moving[1][2], N)
    assert so.rel_rms(ym[1], sm.convolve_moving_receiver(moving[1][0], moving[1][1], idx2, w2)) < 2e-6   # aligned vs grid blocking
    assert np.array_equal(ys[0], sm.convolve_fixed_receiver(static[0][0][None], static[0][1]))


def test_cfg4_long_rir_full_size_properties():
    """configs[3]: 4 channels, P = 120, L = 32768 (8 partitions), N = 2.88 M samples (60 s @ 48 kHz)."""
    from sonicsim_b200 import SonicSim_moving as sm
    rng = np.random.default_rng(4000)
    N, C, L, P = 2880000, 4, 32768, 120
    x, pos = so.synth_dry(rng, N), so.synth_path(rng, P)
    h1 = so.synth_rirs(rng, P, C, L, sr=48000, t60=1.5)
    np.random.seed(4000)
    idx, w = so.setup_dynamic_interp(pos, N)
    y1 = sm.convolve_moving_receiver(x, h1, idx, w)
    assert y1.shape == (C, N) and np.isfinite(y1).all()
    # oracle on the first 2 s only would not see the long tail; use exact properties instead
    h2 = np.zeros_like(h1)
    h2[:, :, 0] = 0.5                                                # delta * 0.5 at every position
    y12 = sm.convolve_moving_receiver(x, (h1 + h2).astype(np.float32), idx, w)
    assert so.rel_rms(y12 - y1, 0.5 * np.broadcast_to(x, (C, N))) < 1e-4     # linearity + delta identity
    # a late single tap (in the last partition) is a pure delay
    h3 = np.zeros_like(h1)
    h3[:, :, L - 7] = 1.0
    y3 = sm.convolve_moving_receiver(x, h3, idx, w)
    ref = np.zeros((C, N), np.float32)
    ref[:, L - 7:] = x[: N - (L - 7)]
    assert so.rel_rms(y3, ref) < 2e-6
    # a slice against the float64 ground truth (first 3 positions' worth of samples)
    n_cut = int(np.searchsorted(idx, 2))
    ref64 = so.convolve_moving_exact_f64(x[:n_cut], h1, idx[:n_cut], w[:n_cut])
    assert so.rel_rms(y1[:, :n_cut], ref64) < TOL
    # ... and three more windows of the full-length render against the same float64 ground truth: one in the middle
    # that straddles a waypoint, one a third of the way in, and the very end.  y[n] only depends on x[n - L + 1 .. n],
    # so a window is evaluated from a slice that starts L - 1 samples earlier and those first samples are dropped.
    mid_wp = int(np.searchsorted(idx, P // 2))                       # first sample of segment P // 2
    for n0, n1 in [(mid_wp - 20000, mid_wp + 20000), (N // 3, N // 3 + 30000), (N - 40000, N)]:
        a = n0 - (L - 1)
        ref_w = so.convolve_moving_exact_f64(x[a:n1], h1, idx[a:n1], w[a:n1])[:, L - 1:]
        assert ref_w.shape == (C, n1 - n0)
        assert so.rel_rms(y1[:, n0:n1], ref_w) < TOL, (n0, n1)


def test_cfg5_binaural_dual_render():
    """configs[4]: 1 moving speaker + noise + music, binaural, plus a second ('direct') RIR set.  The
    reference has no direct-only render; the direct set is the full RIR zeroed 2.5 ms after its
    first peak, rendered as extra sources in the same call."""
    from sonicsim_b200 import render
    rng = np.random.default_rng(5000)
    N, C, L, P, sr = 960000, 2, 4096, 40, 16000
    dry, pos = so.synth_dry(rng, N), so.synth_path(rng, P)
    h = so.synth_rirs(rng, P, C, L)
    peak = np.abs(h).argmax(axis=-1)
    hd = h.copy()
    for p in range(P):
        for c in range(C):
            hd[p, c, peak[p, c] + int(0.0025 * sr):] = 0
    noise, music = so.synth_dry(rng, N), so.synth_dry(rng, N)
    hn, hm = so.synth_rirs(rng, 1, C, L)[0], so.synth_rirs(rng, 1, C, L)[0]
    np.random.seed(5000)
    bounds = render.trajectory_bounds(pos, N)
    R = render.default_renderer()
    outs = R.render_host([render.MovingSource(dry, h, bounds), render.MovingSource(dry, hd, bounds),
                          render.StaticSource(noise, hn), render.StaticSource(music, hm)])
    idx = np.repeat(np.arange(P - 1), np.diff(bounds))
    w = np.concatenate([np.linspace(0, 1, n, endpoint=False) for n in np.diff(bounds)]).astype(np.float32)
    assert so.rel_rms(outs[0], so.convolve_moving_receiver(dry, h, idx, w)) < TOL
    assert so.rel_rms(outs[1], so.convolve_moving_receiver(dry, hd, idx, w)) < TOL
    assert so.rel_rms(outs[2], so.convolve_fixed_receiver(noise[None], hn)) < TOL
    assert so.rel_rms(outs[3], so.convolve_fixed_receiver(music[None], hm)) < TOL


def test_long_rir_compact_trajectory_vs_indexed_and_device_plan():
    """Long RIRs (several partitions accumulated in the frequency domain): the compact-trajectory form and the (idx, w)
    form of the same render against the oracle and against each other; 5 channels, many short segments (blocks that need
    three and more positions), L just above one partition; the device plan equals the host path bit for bit."""
    import torch
    from sonicsim_b200 import SonicSim_moving as sm, render
    R = render.default_renderer()
    for seed, (P, C, L, N) in enumerate([(12, 5, 20000, 150000), (40, 2, 9000, 60000), (3, 1, 4097, 30000)]):
        rng = np.random.default_rng(70 + seed)
        x, h, pos = so.synth_dry(rng, N), so.synth_rirs(rng, P, C, L, sr=48000, t60=1.5), so.synth_path(rng, P)
        np.random.seed(70 + seed)
        idx, w = so.setup_dynamic_interp(pos, N)
        bounds = np.concatenate([[0], np.cumsum(np.bincount(idx, minlength=P - 1))]).astype(np.int32)
        ref = so.convolve_moving_receiver(x, h, idx, w)
        y_b = R.render_host([render.MovingSource(x, h, bounds)])[0]
        y_i = sm.convolve_moving_receiver(x, h, idx, w)
        assert so.rel_rms(y_b, ref) < TOL and so.rel_rms(y_i, ref) < TOL
        assert so.rel_rms(y_b, y_i) < 2e-6
        dev = [render.MovingSource(torch.from_numpy(x).cuda(), torch.from_numpy(h).cuda(), torch.from_numpy(bounds).cuda(), bounds)]
        out = [torch.empty((C, N), device="cuda")]
        plan = R.plan_device(dev, out)
        plan.run(); plan.run()
        torch.cuda.synchronize()
        assert np.array_equal(out[0].cpu().numpy(), y_b)
        plan.close()
