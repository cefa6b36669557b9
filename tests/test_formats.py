"""Formats either side of the path: the WAV writer against the reference's own
output files, the RIR dump reader and the RIR post-processing against the reference's lines."""
import json
import os

import numpy as np
import pytest
import torch

from sonicsim_b200 import formats

REF_WAV = "/root/reference/files/61-908-7127/moving_audio_1.wav"


@pytest.mark.skipif(not os.path.isfile(REF_WAV), reason="/root/reference not present (GPU box)")
def test_wav_writer_is_byte_identical_to_the_references_own_output(tmp_path):
    stem, sr = formats.read_wav_f32(REF_WAV)
    assert stem.shape == (2, 960000) and sr == 16000 and stem.dtype == np.float32
    assert formats.wav_f32_bytes(stem, sr) == open(REF_WAV, "rb").read()


def test_wav_header_layout_and_roundtrip(tmp_path):
    rng = np.random.default_rng(0)
    stem = rng.standard_normal((6, 1000)).astype(np.float32)
    b = formats.wav_f32_bytes(stem, 16000)
    # the layout of the reference's files: RIFF | WAVE | fmt (18 bytes, tag 3) | fact (frames) | data
    assert b[:4] == b"RIFF" and b[8:16] == b"WAVEfmt " and b[16:20] == (18).to_bytes(4, "little")
    assert b[20:22] == (3).to_bytes(2, "little") and b[38:42] == b"fact" and b[46:50] == (1000).to_bytes(4, "little")
    assert b[50:54] == b"data" and len(b) == 58 + 6 * 1000 * 4
    p = tmp_path / "x.wav"
    formats.write_wav_f32(str(p), stem, 16000)
    back, sr = formats.read_wav_f32(str(p))
    assert sr == 16000 and np.array_equal(back, stem)


def test_combine_rirs_matches_reference_lines():
    """SonicSim_audio.py:391-398 restated with torch exactly as written there."""
    rng = np.random.default_rng(1)
    irs = [torch.from_numpy(rng.standard_normal((2, 3000 + 17 * i)).astype(np.float32)) for i in range(5)]
    ref_list = [a[..., :min(x.shape[-1] for x in irs)] for a in irs]                  # clip_all
    ref = torch.stack(ref_list).reshape(5, 1, 2, -1)
    ref = ref / ref.abs().max()
    got = formats.combine_rirs(irs, 5, 1)
    assert got.shape == (5, 1, 2, 3000) and got.dtype == np.float32
    assert np.array_equal(got, ref.numpy())
    assert np.abs(got).max() == 1.0


def test_rir_dump_roundtrip_and_scene_writer(tmp_path):
    rng = np.random.default_rng(2)
    dump = [torch.from_numpy(rng.standard_normal((4, 1, 2, 100)).astype(np.float32)) for _ in range(3)]
    p = tmp_path / "rir_save_train_Binaural.pt"
    torch.save(dump, str(p))                                                          # SonicSet.py:68
    rirs = formats.load_rir_dump(str(p))
    assert len(rirs) == 3 and rirs[0].shape == (4, 2, 100) and np.array_equal(rirs[1], dump[1].numpy()[:, 0])
    w = formats.SceneWriter(workers=2)
    stems = [rng.standard_normal((2, 500)).astype(np.float32) for _ in range(5)]
    side = formats.scene_json([dict(audio=["a.flac"], start_end_points=[(1, 2)], words=["HI"])] * 3,
                              noise=dict(audio=["n.wav"], start_end_points=[(0, 5)]),
                              music=dict(audio=["m.mp3"], start_end_points=[(0, 5)]))
    formats.save_scene(w, str(tmp_path / "scene"), stems[:3], stems[3:], 16000, side)
    w.close()
    names = sorted(os.listdir(tmp_path / "scene"))
    assert names == ["json_data.json", "moving_audio_1.wav", "moving_audio_2.wav", "moving_audio_3.wav",
                     "music_audio.wav", "noise_audio.wav"]
    back, _ = formats.read_wav_f32(str(tmp_path / "scene" / "music_audio.wav"))
    assert np.array_equal(back, stems[4])
    js = json.load(open(tmp_path / "scene" / "json_data.json"))
    assert sorted(js) == ["music", "noise", "source1", "source2", "source3"] and js["source2"]["words"] == ["HI"]


def test_rir_postprocessing_against_golden_of_the_reference_lines(golden):
    """combine_rirs / stack_rirs vs the golden written by exec'ing SonicSim_rir.py:24-41 + SonicSim_audio.py:391-398."""
    g = golden("rir_combine")
    for k in range(int(g["n_cases"])):
        raw = [g[f"raw{k}"][i, :, :l] for i, l in enumerate(g[f"lens{k}"])]
        out = g[f"out{k}"]
        assert np.array_equal(formats.combine_rirs(raw, len(raw), 1), out)
        st = formats.stack_rirs(raw, len(raw), 1)
        assert st.shape == out.shape and np.array_equal(st / np.abs(st).max(), out)
