#!/bin/bash
# ncu capture of the long-RIR render kernel (L = 8192, 2 partitions): bash profiles/prof_long.sh <tag>
TAG=${1:-long}
mkdir -p gpurun_out
cat > /tmp/long_case.py <<'PY'
import sys, os
import numpy as np, torch
sys.path.insert(0, os.getcwd())
from oracle import sonicsim_oracle as so
from sonicsim_b200 import render
R = render.Renderer(0); dev = torch.device("cuda", 0)
rng = np.random.default_rng(0); np.random.seed(0)
L = int(os.environ.get("LONG_L", "8192"))
srcs, outs = [], []
for _ in range(3):
    N, P, C = 960000, 60, 6
    x, h, pos = so.synth_dry(rng, N), so.synth_rirs(rng, P, C, L, t60=1.0), so.synth_path(rng, P)
    b = render.trajectory_bounds(pos, N)
    srcs.append(render.MovingSource(torch.from_numpy(x).to(dev), torch.from_numpy(h).to(dev), torch.from_numpy(b).to(dev), b))
    outs.append(torch.empty((C, N), device=dev))
for _ in range(3):
    R.render_device(srcs, outs)
torch.cuda.synchronize()
PY
ncu --set full --clock-control none --import-source on -k regex:k_render -s 2 -c 1 -o gpurun_out/prof_render_$TAG python /tmp/long_case.py > gpurun_out/ncu_long.log 2>&1
tail -2 gpurun_out/ncu_long.log
