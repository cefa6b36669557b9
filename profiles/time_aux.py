"""Device-resident time of the kernels around the convolution (round 2): loudness (k_kweight<1..3>, k_loud_gate,
k_loud_scale), mixture assembly (k_mix_*), dry-stream assembly (k_dry_assemble), RIR normalisation (k_rir_absmax).
CUDA events, inputs larger than L2 in total, 3 warm-ups.  `python profiles/time_aux.py [reps]`; run it under ncu with
`-k regex:'k_kweight|k_loud|k_mix|k_dry|k_rir'` for the captures in profiles/r2_aux_*."""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import sonicsim_oracle as so                     # synthetic inputs only
from sonicsim_b200 import _lib, dry, render
from sonicsim_b200.SonicSim_audio import gating_plan

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
R = render.Renderer(0)
lib, ctx = R.lib, R.ctx
dev = torch.device("cuda", 0)
stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def timeit(name, fn, nbytes, note=""):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print("%-64s %8.3f ms  %8.1f GB/s algorithmic  %s" % (name, ms, nbytes / ms / 1e6, note))


# ---- loudness: 32 rendered stems (C, N) = (6 -> first 5 weighted, N = 480000), normalised in place
S, C, N, sr = 32, 5, 480000, 16000
stems = torch.randn((S, C, N), device=dev) * 0.05
brk, blo, bhi = gating_plan(N, float(sr), 0.4)
d_brk, d_lo, d_hi = (torch.from_numpy(a).to(dev) for a in (brk, blo, bhi))
n_e = len(brk) - 1
scr = torch.empty((S, 20 * C * n_e), dtype=torch.float64, device=dev)          # SS_LOUD_SCRATCH_DOUBLES
res = torch.empty((S, 2), dtype=torch.float64, device=dev)
items = (_lib.SsLoudItem * S)()
for i in range(S):
    items[i] = _lib.SsLoudItem(data=stems[i].data_ptr(), out=stems[i].data_ptr(), brk=d_brk.data_ptr(), blk_lo=d_lo.data_ptr(),
                               blk_hi=d_hi.data_ptr(), scratch=scr[i].data_ptr(), result=res[i].data_ptr(), stride_n=1,
                               stride_c=N, N=N, C=C, n_e=n_e, n_blocks=len(blo), rate=float(sr), block_size=0.4,
                               target_lufs=-17.0)
timeit("lufs_norm: 32 stems x 5 ch x 30 s (measure + scale in place)",
       lambda: _lib.check(lib.ss_loudness_dev(ctx, items, S, stream)), 4.0 * S * C * N * 2,
       "(read + write of every stem; the three K-weighting passes re-read it from L2)")

# ---- mixture assembly: 16 utterances x (2 speakers + 1 noise), 6 ch x 30 s
U, Sp, M, Cm = 16, 2, 1, 6
E = Cm * N
spk = torch.randn((U, Sp, E), device=dev) * 0.05
noi = torch.randn((U, M, E), device=dev) * 0.02
mix = torch.empty((U, E), device=dev)
sirs = torch.zeros((U, 1), device=dev)
nscr = int(lib.ss_mix_scratch_doubles())
mscr = torch.empty((U, nscr), dtype=torch.float64, device=dev)
mitems = (_lib.SsMixItem * U)()
for u in range(U):
    mitems[u] = _lib.SsMixItem(speakers=spk[u].data_ptr(), noises=noi[u].data_ptr(), sirs=sirs[u].data_ptr(), mix=mix[u].data_ptr(),
                               speakers_out=spk[u].data_ptr(), scratch=mscr[u].data_ptr(), E=E, S=Sp, M=M, snr=15.0)
timeit("mix_stems: 16 utterances x (2 spk + 1 noise) x 6 ch x 30 s",
       lambda: _lib.check(lib.ss_mix_dev(ctx, mitems, U, stream)), 4.0 * U * E * (Sp + M + 1 + Sp),
       "(stems read once, mix + scaled speakers written)")

# ---- dry-stream assembly: 60 s stream from 44.1 kHz stereo clips (resample + mean + place)
total = 60 * sr
clips = [torch.randn((2, 44100 * 11), device=dev) * 0.1 for _ in range(5)]
kt, o, nw, width = dry._resample_kernel(44100, sr, dev)
out = torch.empty((1, total), device=dev)
dc = (_lib.SsDryClip * 5)()
pos = 0
for i, cl in enumerate(clips):
    n = -((-nw * cl.shape[1]) // o)
    dc[i] = _lib.SsDryClip(src=cl.data_ptr(), kernel_t=kt.data_ptr(), dst_start=pos, src_start=0, count=n, channels=2,
                           src_len=cl.shape[1], orig=o, new_rate=nw, width=width, taps=2 * width + o)
    pos += n + 8000
timeit("dry assembly: 60 s @16 kHz from 5 x 11 s stereo 44.1 kHz clips",
       lambda: _lib.check(lib.ss_dry_assemble_dev(ctx, dc, 5, ctypes.c_void_p(out.data_ptr()), total, stream)),
       4.0 * (sum(c.numel() for c in clips) + total), "(475-tap polyphase filter per output sample: compute-bound)")

# ---- RIR normalisation fused into the spectra kernel: cfg2 source with and without SS_RIR_NORMALIZE
rng = np.random.default_rng(0)
P, Cr, L = 40, 6, 4096
np.random.seed(0)
srcs, outs = [], []
for i in range(7):
    x, h = so.synth_dry(rng, N), so.synth_rirs(rng, P, Cr, L)
    b = render.trajectory_bounds(so.synth_path(rng, P), N)
    srcs.append((torch.from_numpy(x).to(dev), torch.from_numpy(h).to(dev), torch.from_numpy(b).to(dev), b))
    outs.append(torch.empty((Cr, N), device=dev))
for flag in (False, True):
    plan = R.plan_device([render.MovingSource(a, b, c, d, flag) for a, b, c, d in srcs], outs)
    timeit("render 7 cfg2 sources, normalize_rirs=%s" % flag, plan.run, 7 * 4.0 * (N + P * Cr * L + Cr * N))
    plan.close()
