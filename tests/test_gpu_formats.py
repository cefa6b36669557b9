"""SURVEY 8f rows 3 and 4 on the device: RIR post-processing fused into the spectra kernel, and the whole
dump -> render -> files pipeline against the oracle pipeline."""
import json
import os

import numpy as np
import pytest
import torch

from conftest import TOL
from oracle import sonicsim_oracle as so

pytestmark = pytest.mark.gpu


def test_rir_normalisation_fused_into_the_spectra_kernel(golden):
    """SS_RIR_NORMALIZE: raw (clipped + stacked) simulator RIRs in, `ir_output /= ir_output.abs().max()`
    (SonicSim_audio.py:398) applied by k_rir_absmax + k_prepare.  Bit-identical to rendering the golden, already
    normalised RIRs of the reference's own lines; host path, device path and static sources."""
    from sonicsim_b200 import formats, render
    g = golden("rir_combine")
    R = render.default_renderer()
    rng = np.random.default_rng(5)
    for k in range(int(g["n_cases"])):
        raw = [g[f"raw{k}"][i, :, :l] for i, l in enumerate(g[f"lens{k}"])]
        ref_h = g[f"out{k}"][:, 0]                                    # (P, C, L) as the reference normalised it
        P = len(raw)
        N = 30000
        x = so.synth_dry(rng, N)
        np.random.seed(k)
        bounds = render.trajectory_bounds(so.synth_path(rng, P), N)
        st = formats.stack_rirs(raw, P, 1)[:, 0]
        want = R.render_host([render.MovingSource(x, ref_h, bounds)])[0]
        got = R.render_host([render.MovingSource(x, st, bounds, normalize_rirs=True)])[0]
        assert np.array_equal(got, want)
        idx = np.repeat(np.arange(P - 1), np.diff(bounds))
        w = np.concatenate([np.linspace(0, 1, n, endpoint=False) for n in np.diff(bounds)]).astype(np.float32)
        assert so.rel_rms(got, so.convolve_moving_receiver(x, ref_h, idx, w)) < TOL
        # device path (a plan = CUDA graph with the extra reduction kernel), mixed with an un-normalised source
        dev = [render.MovingSource(torch.from_numpy(x).cuda(), torch.from_numpy(st).cuda(), torch.from_numpy(bounds).cuda(),
                                   bounds, True),
               render.MovingSource(torch.from_numpy(x).cuda(), torch.from_numpy(ref_h).cuda(), torch.from_numpy(bounds).cuda(),
                                   bounds)]
        outs = [torch.empty(want.shape, device="cuda") for _ in dev]
        plan = R.plan_device(dev, outs)
        plan.run()
        plan.run()                                                    # partial maxima need no reset between runs
        torch.cuda.synchronize()
        assert np.array_equal(outs[0].cpu().numpy(), want) and np.array_equal(outs[1].cpu().numpy(), want)
        plan.close()
        # static source (noise / music RIRs are not normalised by the reference, the flag is still honoured)
        hs = st[0]
        ws = R.render_host([render.StaticSource(x, (hs / np.abs(hs).max()).astype(np.float32))])[0]
        gs = R.render_host([render.StaticSource(x, hs, True)])[0]
        assert np.array_equal(gs, ws)


def test_dump_to_stems_to_files_equals_the_oracle_pipeline(tmp_path):
    """SonicSet.process_single from the RIR dump on (SonicSet.py:68, 77-136): rir_save_*.pt -> render_scene (3 moving
    speakers + noise + music, loudness-normalised on the device) -> SceneWriter -> WAV / JSON files, read back and
    compared with the oracle's serial pipeline (convolve -> lufs_norm -> (C, N) float32)."""
    from sonicsim_b200 import formats, render
    rng = np.random.default_rng(77)
    sr, N, C, L = 16000, 16000 * 6, 2, 1800
    Ps = [7, 5, 9]
    dump = [torch.from_numpy(so.synth_rirs(rng, P, C, L)[:, None]) for P in Ps]          # (P, 1, C, L) per speaker
    path = tmp_path / "rir_save_train_Binaural.pt"
    torch.save(dump, str(path))                                                           # SonicSet.py:68
    dries = [so.synth_dry(rng, N) for _ in Ps]
    poss = [so.synth_path(rng, P) for P in Ps]
    noise_x, music_x = so.synth_dry(rng, N), so.synth_dry(rng, N)
    noise_h, music_h = so.synth_rirs(rng, 1, C, L)[0], so.synth_rirs(rng, 1, C, L)[0]

    rirs = formats.load_rir_dump(str(path))
    np.random.seed(9)
    moving, static = render.render_scene([(d, h, p) for d, h, p in zip(dries, rirs, poss)],
                                         [(noise_x, noise_h), (music_x, music_h)], sr=sr,
                                         moving_lufs=-17, static_lufs=[-24, -29])          # SonicSet.py:97-101
    side = formats.scene_json([dict(audio=["s%d.flac" % i], start_end_points=[(0, N)], words=["W"]) for i in range(3)],
                              noise=dict(audio=["n.wav"], start_end_points=[(0, N)]),
                              music=dict(audio=["m.mp3"], start_end_points=[(0, N)]))
    w = formats.SceneWriter()
    out_dir = str(tmp_path / "scene")
    formats.save_scene(w, out_dir, moving, static, sr, side)
    w.close()

    # the oracle pipeline, in the reference's order of RNG draws (trajectories first, then the five loudness targets)
    np.random.seed(9)
    conv = []
    for d, h, p in zip(dries, rirs, poss):
        idx, wgt = so.setup_dynamic_interp(p, N)
        conv.append(so.convolve_moving_receiver(d, h, idx, wgt))
    conv.append(so.convolve_fixed_receiver(noise_x[None], noise_h))
    conv.append(so.convolve_fixed_receiver(music_x[None], music_h))
    want = []
    for y, lufs in zip(conv, [-17, -17, -17, -24, -29]):
        target = np.random.uniform(lufs - 2, lufs + 2)                                    # SonicSim_audio.py:84
        want.append(so.lufs_norm(np.ascontiguousarray(y.T), sr, target)[0].T)
    names = ["moving_audio_1.wav", "moving_audio_2.wav", "moving_audio_3.wav", "noise_audio.wav", "music_audio.wav"]
    assert sorted(os.listdir(out_dir)) == sorted(names + ["json_data.json"])
    for name, ref in zip(names, want):
        got, got_sr = formats.read_wav_f32(os.path.join(out_dir, name))
        assert got_sr == sr and got.shape == ref.shape and got.dtype == np.float32
        assert so.rel_rms(got, ref) < TOL, name
        # torchaudio reads the file like the reference's consumers do
        import torchaudio
        try:
            ta, ta_sr = torchaudio.load(os.path.join(out_dir, name))
            assert ta_sr == sr and np.array_equal(ta.numpy(), got)
        except (RuntimeError, ImportError):
            pass                                                                           # no audio backend in this image
    js = json.load(open(os.path.join(out_dir, "json_data.json")))
    assert sorted(js) == ["music", "noise", "source1", "source2", "source3"]


def test_example_script_renders_a_dump(tmp_path):
    """examples/render_scene_from_dump.py end to end: dump + dry WAVs + positions.npy in, loudness-normalised stems out."""
    import importlib.util
    from sonicsim_b200 import formats
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("render_scene_from_dump", os.path.join(root, "examples", "render_scene_from_dump.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    rng = np.random.default_rng(5)
    sr, N, C, L, Ps = 16000, 16000 * 4, 2, 1200, [5, 6]
    dump = [torch.from_numpy(so.synth_rirs(rng, P, C, L)[:, None]) for P in Ps]
    torch.save(dump, str(tmp_path / "rir_save_train_Binaural.pt"))
    poss = [so.synth_path(rng, P) for P in Ps]
    np.save(str(tmp_path / "positions.npy"), np.array(poss, dtype=object), allow_pickle=True)
    dries = [so.synth_dry(rng, N) for _ in Ps]
    paths = []
    for i, d in enumerate(dries):
        p = str(tmp_path / ("dry%d.wav" % i))
        formats.write_wav_f32(p, d[None], sr)
        paths.append(p)
    out_dir = str(tmp_path / "out")
    np.random.seed(21)
    mod.main(["render_scene_from_dump.py", str(tmp_path / "rir_save_train_Binaural.pt")] + paths + [out_dir])
    np.random.seed(21)
    conv = []
    for d, h, p in zip(dries, dump, poss):
        idx, w = so.setup_dynamic_interp(np.asarray(p, dtype=float), N)
        conv.append(so.convolve_moving_receiver(d, h.numpy()[:, 0], idx, w))
    for i, y in enumerate(conv):
        target = np.random.uniform(-19, -15)
        want = so.lufs_norm(np.ascontiguousarray(y.T), sr, target)[0].T
        got, got_sr = formats.read_wav_f32(os.path.join(out_dir, "moving_audio_%d.wav" % (i + 1)))
        assert got_sr == sr and so.rel_rms(got, want) < TOL
