"""Kernel split of the long-RIR case (configs[3]: 3 moving, 4 ch, 120 pt, L=32768, 60 s @48 kHz) and of
shorter variants; device-resident inputs, per-kernel CUDA events (ss_set_profiling)."""
import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import sonicsim_oracle as so
from sonicsim_b200 import render

R = render.Renderer(0)
dev = torch.device("cuda", 0)


def moving(rng, N, P, C, L, sr, t60):
    x, h, pos = so.synth_dry(rng, N), so.synth_rirs(rng, P, C, L, sr=sr, t60=t60), so.synth_path(rng, P)
    b = render.trajectory_bounds(pos, N)
    return render.MovingSource(torch.from_numpy(x).to(dev), torch.from_numpy(h).to(dev), torch.from_numpy(b).to(dev), b), \
        torch.empty((C, N), device=dev)


rng = np.random.default_rng(0)
np.random.seed(0)
for name, (n_src, N, P, C, L, sr) in {"L=32768 48k": (3, 2880000, 120, 4, 32768, 48000),
                                      "L=16384 16k": (3, 960000, 60, 6, 16384, 16000),
                                      "L=8192 16k": (3, 960000, 60, 6, 8192, 16000)}.items():
    trip = [moving(rng, N, P, C, L, sr, 1.5) for _ in range(n_src)]
    srcs, outs = [p[0] for p in trip], [p[1] for p in trip]
    for _ in range(2):
        R.render_device(srcs, outs)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        R.render_device(srcs, outs)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    R.set_profiling(True)
    R.render_device(srcs, outs); torch.cuda.synchronize()
    a, b, n = R.get_profile()
    R.set_profiling(False)
    print("%-14s step %.3f ms | profiled (serial): k_prepare %.3f ms, k_render %.3f ms, %d launch pairs" % (name, ms, a, b, n))
    del trip, srcs, outs
    torch.cuda.empty_cache()
