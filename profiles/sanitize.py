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
# round 2: several items per CTA in k_render_fast (item ring, X kept / replaced), a device plan (CUDA graph),
# on-device RIR normalisation, dry assembly with resampling, loudness fused behind a host-path render
from sonicsim_b200 import dry
xb = so.synth_dry(rng, 140000)
hb = so.synth_rirs(rng, 9, 2, 900)
np.random.seed(2)
bb = render.trajectory_bounds(so.synth_path(rng, 9), 140000)
R = render.default_renderer()
yb = R.render_host([render.MovingSource(xb, hb * 3.0, bb, None, True)], lufs_targets=[-20.0])[0]
dsrc = [render.MovingSource(torch.from_numpy(xb).cuda(), torch.from_numpy(hb).cuda(), torch.from_numpy(bb).cuda(), bb)]
dout = [torch.empty((2, 140000), device="cuda")]
plan = R.plan_device(dsrc, dout)
plan.run(); plan.run(); torch.cuda.synchronize(); plan.close()
import ctypes
from sonicsim_b200 import _lib
kt, o, nw, width = dry._resample_kernel(44100, 16000, torch.device("cuda"))
cl = torch.randn((2, 44100), device="cuda")
outd = torch.empty((1, 20000), device="cuda")
dc = (_lib.SsDryClip * 1)(_lib.SsDryClip(src=cl.data_ptr(), kernel_t=kt.data_ptr(), dst_start=100, src_start=0, count=16000, channels=2,
                                          src_len=44100, orig=o, new_rate=nw, width=width, taps=2 * width + o))
_lib.check(R.lib.ss_dry_assemble_dev(R.ctx, dc, 1, ctypes.c_void_p(outd.data_ptr()), 20000, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
torch.cuda.synchronize()
ref = so.convolve_moving_receiver(x, h, idx, w)
print("ok", so.rel_rms(y, ref), so.rel_rms(y2.numpy(), ref), float(g))
