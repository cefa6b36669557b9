"""Generate tests/golden/*.npz by running the UNMODIFIED reference functions
(/root/reference/SonicSim-SonicSet/SonicSim_moving.py:15-125, SonicSim_audio.py:17-47)
on seeded synthetic inputs.  Run in the authoring container only:

    python oracle/make_golden.py

The fixtures hold both the inputs and the reference's outputs, so the tests that read
them need neither /root/reference nor this script.  TEST INFRASTRUCTURE.
"""
import os
import sys
import types

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


def golden_mix_noisy():
    """f1 (enhancement variant): overlap_audio (:34-48), find_overlap_region (:50-75) and the noisy-mixture block of
    the enhancement dataloader (enhancement/look2hear/datas/movingdatamodule.py235-257), exec'd verbatim."""
    import random
    import textwrap
    import torch
    src = open("/root/reference/enhancement/look2hear/datas/movingdatamodule.py").read().splitlines()
    assert src[28].startswith("def compute_mch_rms_dB") and src[33].startswith("def overlap_audio"), (src[28], src[33])
    assert src[49].startswith("def find_overlap_region") and src[75].strip() == "", (src[49], src[75])
    ns = {"torch": torch, "np": np, "random": random}
    exec("\n".join(src[28:75]), ns)
    assert src[234].strip().startswith("all_noise = torch.sum") and src[256].strip().startswith("mix_wav ="), (src[234], src[256])
    body = textwrap.dedent("\n".join(src[234:257]))
    cases = {}
    for k, (seed, T, rows, sr, delay) in enumerate([(71, 4000, 1, 500, 6), (72, 1000, 2, 16000, 0.01), (73, 300, 1, 100, 6)]):
        rng = np.random.default_rng(seed)
        x = rng.standard_normal((rows, T)).astype(np.float32)
        y = ns["overlap_audio"](torch.from_numpy(x.copy()), sr, delay=delay)
        cases[f"ov_x{k}"], cases[f"ov_y{k}"] = x, y.numpy()
        cases[f"ov_sr{k}"], cases[f"ov_delay{k}"] = np.int64(sr), np.float64(delay)
    cases["n_ov"] = np.int64(3)
    for k, (seed, kw) in enumerate([(81, {}), (82, dict(min_overlap=1, max_overlap=2)),
                                    (83, dict(max_duration=0.5, sample_rate=1000))]):
        rng = np.random.default_rng(seed)
        data = {}
        for s in range(3):
            starts = np.sort(rng.integers(0, 9000, 4))
            data[f"spk{s}"] = {"start_end_points": [[int(a), int(a) + int(rng.integers(100, 900))] for a in starts]}
        data["noise"] = {"other": 1}
        random.seed(seed)
        lo, hi = ns["find_overlap_region"](data, **kw)
        cases[f"fo_points{k}"] = np.array([data[f"spk{s}"]["start_end_points"] for s in range(3)], dtype=np.int64)
        cases[f"fo_out{k}"] = np.array([lo, hi], dtype=np.int64)
        cases[f"fo_seed{k}"] = np.int64(seed)
        cases[f"fo_kw{k}"] = np.array(repr(kw))
    cases["n_fo"] = np.int64(3)
    for k, (seed, M, T, sr, kinds) in enumerate([(91, 2, 12000, 1000, ["music", "noise"]), (92, 1, 8000, 1000, ["noise"]),
                                                 (93, 1, 4000, 16000, ["music"])]):
        rng = np.random.default_rng(seed)
        spk = (rng.standard_normal(T) * 0.1).astype(np.float32)
        noi = (rng.standard_normal((M, T)) * (1e-5 if k == 2 else 0.05)).astype(np.float32)
        torch.manual_seed(seed)
        snr = torch.Tensor(1).uniform_(-10, 15).numpy()
        torch.manual_seed(seed)
        env = dict(ns)
        env.update(self=types.SimpleNamespace(sample_rate=sr), speaker_wavs=torch.from_numpy(spk.copy()),
                   noise_wav=torch.from_numpy(noi.copy()), noise_types=kinds)
        exec(body, env)
        cases[f"mn_spk{k}"], cases[f"mn_noise{k}"], cases[f"mn_snr{k}"], cases[f"mn_sr{k}"] = spk, noi, snr, np.int64(sr)
        cases[f"mn_mix{k}"] = env["mix_wav"].numpy()
    cases["n_mn"] = np.int64(3)
    np.savez_compressed(os.path.join(OUT, "mix_noisy.npz"), **cases)
    print("mix_noisy.npz", os.path.getsize(os.path.join(OUT, "mix_noisy.npz")))


def golden_dry():
    """f2: dry-stream assembly - the unmodified create_long_audio / create_background_audio
    (SonicSim_audio.py:231-340) on the synthetic directory of oracle/dry_fixture.py, torchaudio.load stubbed."""
    import json
    import random
    import tempfile
    import pathlib
    import torchaudio
    from oracle import dry_fixture
    _, ref = ref_loader.load(want_audio=True)
    cases = []
    with tempfile.TemporaryDirectory() as tmp:
        spk, noise_json, load = dry_fixture.build(pathlib.Path(tmp))
        ref.torchaudio = types.SimpleNamespace(transforms=torchaudio.transforms, load=load)
        real_walk = os.walk
        os.walk = dry_fixture.sorted_walk(real_walk)
        try:
            ref.print("")      # rich's first print draws from `random`
            for seed in range(6):
                random.seed(seed)
                a, se, names = ref.create_long_audio(spk, 60)
                random.seed(100 + seed)
                b, bse, bnames = ref.create_background_audio(noise_json, 60)
                cases.append({"seed": seed, "speech_sha256": dry_fixture.digest(a), "speech_shape": list(a.shape),
                              "speech_spans": [list(map(int, x)) for x in se],
                              "speech_names": [os.path.basename(n) for n in names],
                              "bg_sha256": dry_fixture.digest(b), "bg_shape": list(b.shape),
                              "bg_spans": [list(map(int, x)) for x in bse],
                              "bg_names": [os.path.basename(n) for n in bnames]})
        finally:
            os.walk = real_walk
    with open(os.path.join(OUT, "dry_assembly.json"), "w") as f:
        json.dump({"cases": cases}, f, indent=1)
    print("dry_assembly.json", len(cases), "cases")


def golden_rir_combine():
    """f4: RIR post-processing - the reference's own lines, exec'd verbatim: clip_all (SonicSim_rir.py:24-41; the module
    needs habitat_sim, so the function's source lines are exec'd) and the tail of generate_rir_combination
    (SonicSim_audio.py:391-398: clip, stack, reshape, divide by the global abs-max)."""
    import textwrap
    import torch
    rir_src = open("/root/reference/SonicSim-SonicSet/SonicSim_rir.py").read().splitlines()
    aud_src = open("/root/reference/SonicSim-SonicSet/SonicSim_audio.py").read().splitlines()
    assert rir_src[23].startswith("def clip_all") and rir_src[40].strip() == "return clipped_audio_list", (rir_src[23], rir_src[40])
    assert aud_src[390].strip().startswith("ir_list = clip_all(ir_list)") and aud_src[397].strip().startswith("ir_output /="), \
        (aud_src[390], aud_src[397])
    ns = {"torch": torch}
    exec("\n".join(rir_src[23:41]), ns)
    body = textwrap.dedent("\n".join(aud_src[390:398]))
    cases = {}
    for k, (seed, P, C, L0) in enumerate([(81, 5, 2, 700), (82, 6, 3, 2100), (83, 3, 1, 33)]):
        rng = np.random.default_rng(seed)
        lens = [L0 + int(rng.integers(0, 40)) for _ in range(P)]
        raw = [(so.synth_rirs(rng, 1, C, l)[0] * float(rng.uniform(0.2, 3.0))).astype(np.float32) for l in lens]
        env = dict(ns)
        env.update(ir_list=[torch.from_numpy(r.copy()) for r in raw], source_idx_list=list(range(P)), receiver_idx_list=[0])
        exec(body, env)
        out = env["ir_output"].numpy()
        assert out.shape == (P, 1, C, min(lens)) and out.dtype == np.float32
        pad = np.zeros((P, C, max(lens)), np.float32)
        for i, r in enumerate(raw):
            pad[i, :, :lens[i]] = r
        cases[f"raw{k}"], cases[f"lens{k}"], cases[f"out{k}"] = pad, np.asarray(lens, np.int64), out
    cases["n_cases"] = np.int64(3)
    np.savez_compressed(os.path.join(OUT, "rir_combine.npz"), **cases)
    print("rir_combine.npz", 3, "cases")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "rir_combine":
        golden_rir_combine()
    elif len(sys.argv) > 1 and sys.argv[1] == "mix_noisy":
        golden_mix_noisy()
    elif len(sys.argv) > 1 and sys.argv[1] == "dry":
        golden_dry()
    else:
        main()
        golden_mix_noisy()
        golden_dry()
        golden_rir_combine()
