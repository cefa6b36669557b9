"""The call pattern of SonicSet.process_single (SonicSet.py:77-101) against the drop-in modules installed
under the reference's module names, and against the oracle doing the same calls."""
import sys

import numpy as np
import pytest
import torch

from conftest import TOL
from oracle import sonicsim_oracle as so

pytestmark = pytest.mark.gpu


def test_process_single_call_pattern():
    import sonicsim_b200
    sonicsim_b200.install_dropin()
    import SonicSim_audio          # noqa: E402  (what SonicSet.py:19-20 imports)
    import SonicSim_moving         # noqa: E402
    assert SonicSim_moving is sys.modules["sonicsim_b200.SonicSim_moving"]

    rng = np.random.default_rng(7)
    sample_rate, N, C, L, P = 16000, 96000, 2, 2000, 6
    # what generate_rir_combination returns per speaker: (P, 1, C, L) CPU tensors (SonicSim_audio.py:397)
    ir_outputs = [torch.from_numpy(so.synth_rirs(rng, P, C, L)[:, None]) for _ in range(3)]
    nav_points = [[list(p) for p in so.synth_path(rng, P)] for _ in range(3)]
    dry = [torch.from_numpy(so.synth_dry(rng, N)[None]) for _ in range(3)]          # create_long_audio: (1, N)
    noise_audio, music_audio = torch.from_numpy(so.synth_dry(rng, N)[None]), torch.from_numpy(so.synth_dry(rng, N)[None])
    rir_noise, rir_music = torch.from_numpy(so.synth_rirs(rng, 1, C, L)[0]), torch.from_numpy(so.synth_rirs(rng, 1, C, L)[0])

    def pipeline(moving_mod, audio_mod, fixed):
        np.random.seed(11)
        recv = [moving_mod(dry[i], ir_outputs[i], nav_points[i]) for i in range(3)]                      # :77-79
        n = torch.from_numpy(fixed(noise_audio, rir_noise.cpu()))                                       # :93
        m = torch.from_numpy(fixed(music_audio, rir_music.cpu()))                                       # :94
        outs = [audio_mod(r.transpose(0, 1).numpy(), sample_rate, -17)[0] for r in recv]                # :97-99
        outs.append(audio_mod(n.transpose(0, 1).numpy(), sample_rate, -24)[0])                          # :100
        outs.append(audio_mod(m.transpose(0, 1).numpy(), sample_rate, -29)[0])                          # :101
        return [torch.from_numpy(np.ascontiguousarray(o)).transpose(0, 1) for o in outs]                # as saved at :102-106

    ours = pipeline(SonicSim_moving.interpolate_moving_audio, SonicSim_audio.get_lufs_norm_audio,
                    SonicSim_moving.convolve_fixed_receiver)
    ref = pipeline(lambda a, b, c: torch.from_numpy(so.interpolate_moving_audio(a.numpy(), b.numpy(), c)),
                   so.get_lufs_norm_audio, so.convolve_fixed_receiver)
    for a, b in zip(ours, ref):
        assert a.shape == (C, N) and a.dtype == torch.float32
        assert so.rel_rms(a.numpy(), b.numpy()) < TOL


def test_numpy_semantics_at_the_boundary():
    """What NumPy / SciPy do for the reference at SonicSim_moving.py:86-94 and what the drop-in must therefore do too:
    negative interp_index wraps (idx = -1 pairs the LAST position with the FIRST), non-integer index arrays raise
    IndexError, float64 inputs give a float64 result (computed in float32 here), more RIRs than waypoints is fine."""
    from sonicsim_b200 import SonicSim_moving as sm
    rng = np.random.default_rng(17)
    P, C, L, N = 6, 2, 700, 20000
    x, h = so.synth_dry(rng, N), so.synth_rirs(rng, P, C, L)
    w = rng.random(N).astype(np.float32)
    # negative indices, with and without the -1 wrap pair
    for lo in (-1, -2):
        idx = rng.integers(-P, lo + 1, N)
        idx[::3] = rng.integers(0, P - 1, len(idx[::3]))
        ref = so.convolve_moving_receiver(x, h, idx, w)
        got = sm.convolve_moving_receiver(x, h, idx, w)
        assert got.dtype == np.float32 and so.rel_rms(got, ref) < TOL
    with pytest.raises(IndexError):
        sm.convolve_moving_receiver(x, h, np.full(N, -P - 1), w)
    with pytest.raises(IndexError):
        so.convolve_moving_receiver(x, h, np.full(N, -P - 1), w)
    with pytest.raises(IndexError):
        sm.convolve_moving_receiver(x, h, np.zeros(N, np.float32), w)                 # float index array
    with pytest.raises(IndexError):
        so.convolve_moving_receiver(x, h, np.zeros(N, np.float32), w)
    # dtype rule
    idx = np.sort(rng.integers(0, P - 1, N))
    ref64 = so.convolve_moving_receiver(x.astype(np.float64), h.astype(np.float64), idx, w.astype(np.float64))
    got64 = sm.convolve_moving_receiver(x.astype(np.float64), h, idx, w)
    assert ref64.dtype == np.float64 and got64.dtype == np.float64 and so.rel_rms(got64, ref64) < 2e-6
    assert sm.convolve_moving_receiver(x, h, idx, w.astype(np.float64)).dtype == \
        so.convolve_moving_receiver(x, h, idx, w.astype(np.float64)).dtype
    f64 = sm.convolve_fixed_receiver(x[None].astype(np.float64), h[0])
    assert f64.dtype == so.convolve_fixed_receiver(x[None].astype(np.float64), h[0]).dtype == np.float64
    # interpolate_moving_audio: RIR count vs waypoint count
    pos = so.synth_path(rng, P - 2)
    np.random.seed(4)
    a = sm.interpolate_moving_audio(torch.from_numpy(x[None]), torch.from_numpy(h[:, None]), pos)          # 6 RIRs, 4 waypoints
    np.random.seed(4)
    b = so.interpolate_moving_audio(x[None], h[:, None], pos)
    assert so.rel_rms(a.numpy(), b) < TOL
    pos = so.synth_path(rng, P + 2)
    with pytest.raises(IndexError):
        np.random.seed(4)
        sm.interpolate_moving_audio(torch.from_numpy(x[None]), torch.from_numpy(h[:, None]), pos)          # 6 RIRs, 8 waypoints
    with pytest.raises(IndexError):
        np.random.seed(4)
        so.interpolate_moving_audio(x[None], h[:, None], pos)
