"""Generate tests/golden/*.npz by running the UNMODIFIED reference functions
(/root/reference/SonicSim-SonicSet/SonicSim_moving.py:15-125, SonicSim_audio.py:17-47)
on seeded synthetic inputs.  Run in the authoring container only:

    python oracle/make_golden.py

The fixtures hold both the inputs and the reference's outputs, so the tests that read
them need neither /root/reference nor this script.  TEST INFRASTRUCTURE.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_loader                      # noqa: E402
from oracle import sonicsim_oracle as so           # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def main():
    import torch
    ref, ref_audio = ref_loader.load(want_audio=True)
    os.makedirs(OUT, exist_ok=True)

    # ---- a1: setup_dynamic_interp (seed np.random right before the call)
    cases = {}
    for k, (seed, P, N) in enumerate([(11, 5, 1000), (12, 40, 48000), (13, 3, 7), (14, 9, 5),
                                      (15, 2, 333)]):
        rng = np.random.default_rng(seed)
        pos = so.synth_path(rng, P)
        if k == 1:
            pos[7] = pos[6]                         # duplicate waypoint -> zero-length segment
        np.random.seed(seed)
        idx, w = ref.setup_dynamic_interp(pos, N)
        cases[f"pos{k}"] = pos
        cases[f"N{k}"] = np.int64(N)
        cases[f"seed{k}"] = np.int64(seed)
        cases[f"idx{k}"] = idx.astype(np.int32)
        cases[f"w{k}"] = w
    cases["n_cases"] = np.int64(5)
    np.savez_compressed(os.path.join(OUT, "setup_dynamic_interp.npz"), **cases)

    # ---- a4: convolve_fixed_receiver
    cases = {}
    for k, (seed, C, L, N) in enumerate([(21, 1, 300, 2000), (22, 2, 1025, 9000), (23, 3, 64, 50),
                                         (24, 2, 5000, 4500)]):
        rng = np.random.default_rng(seed)
        x = so.synth_dry(rng, N).reshape(1, N)
        h = so.synth_rirs(rng, 1, C, L)[0]
        y = ref.convolve_fixed_receiver(x, h)
        y_t = ref.convolve_fixed_receiver(torch.from_numpy(x), torch.from_numpy(h))   # SonicSet.py:93 passes tensors
        assert np.array_equal(np.asarray(y), np.asarray(y_t))
        cases[f"x{k}"], cases[f"h{k}"], cases[f"y{k}"] = x, h, np.asarray(y, dtype=np.float32)
        assert y.dtype == np.float32
    cases["n_cases"] = np.int64(4)
    np.savez_compressed(os.path.join(OUT, "convolve_fixed_receiver.npz"), **cases)

    # ---- a2: convolve_moving_receiver with the reference's own (idx, w)
    cases = {}
    for k, (seed, P, C, L, N) in enumerate([(31, 5, 2, 257, 6000), (32, 2, 1, 100, 1500),
                                            (33, 7, 3, 2100, 4000),      # L >= N/2: oaconvolve falls back to fftconvolve
                                            (34, 12, 2, 4500, 20000),    # L > 4096: two partitions in the CUDA path
                                            (35, 6, 2, 31, 9)]):         # N < typical block, tiny
        rng = np.random.default_rng(seed)
        x = so.synth_dry(rng, N)
        h = so.synth_rirs(rng, P, C, L)
        pos = so.synth_path(rng, P)
        np.random.seed(seed)
        idx, w = ref.setup_dynamic_interp(pos, N)
        y = ref.convolve_moving_receiver(x, h, idx, w)
        assert y.dtype == np.float32, y.dtype
        cases[f"x{k}"], cases[f"h{k}"] = x, h
        cases[f"idx{k}"], cases[f"w{k}"] = idx.astype(np.int32), w
        cases[f"y{k}"] = y
    cases["n_cases"] = np.int64(5)
    np.savez_compressed(os.path.join(OUT, "convolve_moving_receiver.npz"), **cases)

    # ---- a3: interpolate_moving_audio (torch in, torch out, RNG seeded before the call)
    cases = {}
    for k, (seed, P, C, L, N) in enumerate([(41, 6, 2, 400, 8000), (42, 4, 1, 1000, 5000)]):
        rng = np.random.default_rng(seed)
        x = so.synth_dry(rng, N).reshape(1, N)
        h = so.synth_rirs(rng, P, C, L).reshape(P, 1, C, L)
        pos = so.synth_path(rng, P)
        np.random.seed(seed)
        y = ref.interpolate_moving_audio(torch.from_numpy(x), torch.from_numpy(h), [list(p) for p in pos])
        assert isinstance(y, torch.Tensor) and y.dtype == torch.float32
        cases[f"x{k}"], cases[f"h{k}"], cases[f"pos{k}"] = x, h, pos
        cases[f"seed{k}"] = np.int64(seed)
        cases[f"y{k}"] = y.numpy()
    cases["n_cases"] = np.int64(2)
    np.savez_compressed(os.path.join(OUT, "interpolate_moving_audio.npz"), **cases)

    # ---- a6: fft_conv (torch.fft)
    cases = {}
    for k, (seed, L, N) in enumerate([(51, 200, 3000), (52, 4096, 6000),          # N+L-1 odd: the reference's irfftn returns N+L-2 resampled samples
                                      (53, 200, 3001), (54, 4096, 6001)]):        # N+L-1 even: a true full convolution
        rng = np.random.default_rng(seed)
        x = so.synth_dry(rng, N)
        h = so.synth_rirs(rng, 1, 1, L)[0, 0]
        y = ref_audio.fft_conv(torch.from_numpy(x), torch.from_numpy(h), is_cpu=True)
        cases[f"x{k}"], cases[f"h{k}"], cases[f"y{k}"] = x, h, y.numpy()
    cases["n_cases"] = np.int64(4)
    np.savez_compressed(os.path.join(OUT, "fft_conv.npz"), **cases)

    # ---- f1: mixture assembly - the reference's own lines 105-124 of the dataloader, exec'd verbatim
    # (the module itself cannot be imported: it needs pytorch_lightning / dataset files)
    import textwrap
    import types
    src = open("/root/reference/separation/look2hear/datas/movingdatamodule.py").read().splitlines()
    assert src[28].startswith("def compute_mch_rms_dB") and src[104].strip() == "# Random SIR and SNR", (src[28], src[104])
    ns = {"torch": torch, "np": np}
    exec("\n".join(src[28:32]), ns)                                   # compute_mch_rms_dB, :29-32
    body = textwrap.dedent("\n".join(src[104:124]))                   # :105-124 (through mix_wav = ...)
    cases = {}
    for k, (seed, S, M, C, T) in enumerate([(61, 2, 1, 1, 16000), (62, 3, 2, 2, 4000), (63, 2, 1, 6, 2000)]):
        rng = np.random.default_rng(seed)
        shape_s = (S, T) if C == 1 else (S, C, T)
        shape_n = (M, T) if C == 1 else (M, C, T)
        spk = (rng.standard_normal(shape_s) * rng.uniform(0.01, 0.3, (S,) + (1,) * (len(shape_s) - 1))).astype(np.float32)
        noi = (rng.standard_normal(shape_n) * 0.05).astype(np.float32)
        if k == 2:
            spk[1] *= 1e-4                                           # interferer far below the target: +40 dB clamp
        torch.manual_seed(seed)
        sirs = torch.Tensor(S - 1).uniform_(-6, 6).numpy()           # the draws the exec'd lines will make
        snr = torch.Tensor(1).uniform_(10, 20).numpy()
        torch.manual_seed(seed)
        env = dict(ns)
        env.update(self=types.SimpleNamespace(num_spks=S), speaker_wav=torch.from_numpy(spk.copy()),
                   noise_wav=torch.from_numpy(noi.copy()))
        exec(body, env)
        cases[f"spk{k}"], cases[f"noise{k}"], cases[f"sirs{k}"], cases[f"snr{k}"] = spk, noi, sirs, snr
        cases[f"mix{k}"] = env["mix_wav"].numpy()
        cases[f"spk_out{k}"] = env["speaker_wav"].numpy()
    cases["n_cases"] = np.int64(3)
    np.savez_compressed(os.path.join(OUT, "mix_stems.npz"), **cases)

    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
