"""Small renders for compute-sanitizer (memcheck / racecheck / initcheck)."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import sonicsim_oracle as so
from sonicsim_b200 import SonicSim_moving as sm, SonicSim_audio as sa, render, mix
import torch
rng = np.random.default_rng(0)
x, h, pos = so.synth_dry(rng, 20000), so.synth_rirs(rng, 5, 3, 700), so.synth_path(rng, 5)
np.random.seed(0)
idx, w = so.setup_dynamic_interp(pos, 20000)
y = sm.convolve_moving_receiver(x, h, idx, w)                       # indexed, grid blocking
np.random.seed(0)
y2 = sm.interpolate_moving_audio(torch.from_numpy(x[None]), torch.from_numpy(h[:, None]), pos)   # bounds, aligned, FAST
ys = sm.convolve_fixed_receiver(x[None], h[0])                      # static
hl = so.synth_rirs(rng, 3, 2, 9000)
np.random.seed(1)
i2, w2 = so.setup_dynamic_interp(so.synth_path(rng, 3), 20000)
yl = sm.convolve_moving_receiver(x, hl, i2, w2)                     # LONG variant
n, g = sa.lufs_norm(np.ascontiguousarray(ys.T), 16000, -20.0)       # loudness kernels
m, s = mix.mix_stems(torch.from_numpy(np.stack([y, y * 0.5])), torch.from_numpy(ys[None]), [0.0], 15.0)
ref = so.convolve_moving_receiver(x, h, idx, w)
print("ok", so.rel_rms(y, ref), so.rel_rms(y2.numpy(), ref), float(g))
