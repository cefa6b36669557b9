#!/usr/bin/env python
"""Render one SonicSet scene from a pre-baked RIR dump, without Habitat.

    python examples/render_scene_from_dump.py rir_save_train_Binaural.pt dry1.wav dry2.wav dry3.wav out_dir

The dump is what SonicSet.py:68 writes (list of per-speaker (P, 1, C, L) tensors); the dry files are
float32 mono WAVs (what create_long_audio would have assembled); waypoints are taken as equally spaced
if no `positions.npy` (list of (P, 3) arrays) sits next to the dump.  Stems are loudness-normalised like
SonicSet.py:97-99 and written as float32 WAV like SonicSet.py:102-104.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sonicsim_b200 import formats, render          # noqa: E402


def main(argv):
    dump, dries, out_dir = argv[1], argv[2:-1], argv[-1]
    rirs = formats.load_rir_dump(dump)
    pos_path = os.path.join(os.path.dirname(dump), "positions.npy")
    positions = list(np.load(pos_path, allow_pickle=True)) if os.path.isfile(pos_path) else \
        [np.stack([np.arange(r.shape[0]), np.zeros(r.shape[0]), np.zeros(r.shape[0])], 1).astype(float) for r in rirs]
    moving = []
    for path, h, pos in zip(dries, rirs, positions):
        x, sr = formats.read_wav_f32(path)
        moving.append((x[0], h, pos))
    stems, _ = render.render_scene(moving, [], sr=sr, moving_lufs=-17)
    w = formats.SceneWriter()
    formats.save_scene(w, out_dir, stems, [], sr)
    w.close()
    print("wrote %d stems to %s" % (len(stems), out_dir))


if __name__ == "__main__":
    main(sys.argv)
