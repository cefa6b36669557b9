"""Device-resident render time of the BASELINE.json configs (informational; bench.py's line is configs[1])."""
import sys, os, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import sonicsim_oracle as so
from sonicsim_b200 import render

R = render.Renderer(0)
dev = torch.device("cuda", 0)


def moving(rng, N, P, C, L, sr=16000, t60=0.5):
    x, h, pos = so.synth_dry(rng, N), so.synth_rirs(rng, P, C, L, sr=sr, t60=t60), so.synth_path(rng, P)
    b = render.trajectory_bounds(pos, N)
    return render.MovingSource(torch.from_numpy(x).to(dev), torch.from_numpy(h).to(dev), torch.from_numpy(b).to(dev), b), \
        torch.empty((C, N), device=dev)


def static(rng, N, C, L):
    x, h = so.synth_dry(rng, N), so.synth_rirs(rng, 1, C, L)[0]
    return render.StaticSource(torch.from_numpy(x).to(dev), torch.from_numpy(h).to(dev)), torch.empty((C, N), device=dev)


def timeit(name, srcs, outs, audio_s, reps=10):
    for _ in range(3):
        R.render_device(srcs, outs)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        R.render_device(srcs, outs)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print("%-58s %9.3f ms  %12.0f audio-s/s" % (name, ms, audio_s / (ms / 1e3)))


rng = np.random.default_rng(0)
np.random.seed(0)
s, o = static(rng, 64000, 1, 4096)
timeit("cfg1: 1 static, mono, L=4096, 4 s (single call)", [s], [o], 4.0, 50)
pairs = [moving(rng, 480000, 40, 6, 4096) for _ in range(2)]
timeit("cfg2: 2 spk, 6 mic, 40 pt, 30 s (single utterance)", [p[0] for p in pairs], [p[1] for p in pairs], 30.0, 20)
items = []
for _ in range(8):
    items += [moving(rng, 960000, 60, 6, 4096) for _ in range(2)] + [static(rng, 960000, 6, 4096) for _ in range(2)]
timeit("cfg3: 8 utt x (2 moving 60 pt + 2 static), 6 mic, 60 s", [p[0] for p in items], [p[1] for p in items], 8 * 60.0, 5)
trip = [moving(rng, 2880000, 120, 4, 32768, sr=48000, t60=1.5) for _ in range(3)]
timeit("cfg4: 3 moving, 4 ch, 120 pt, L=32768, 60 s @48k", [p[0] for p in trip], [p[1] for p in trip], 60.0, 3)
enh = [moving(rng, 960000, 40, 2, 4096), moving(rng, 960000, 40, 2, 4096), static(rng, 960000, 2, 4096), static(rng, 960000, 2, 4096)]
timeit("cfg5: 1 moving (+direct set) + 2 static, binaural, 60 s", [p[0] for p in enh], [p[1] for p in enh], 60.0, 20)
