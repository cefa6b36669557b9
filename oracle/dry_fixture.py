"""Synthetic LibriSpeech-like speaker directory + noise-length JSON and a deterministic stand-in for
`torchaudio.load`, shared by oracle/make_golden.py (which runs the unmodified reference on it) and
tests/test_dry.py.  TEST INFRASTRUCTURE."""
import hashlib
import json
import os

import numpy as np


def fake_loader(lengths, stereo=()):
    """`torchaudio.load` replacement: content is a function of the file's basename only."""
    import torch

    def load(path):
        n = lengths[os.path.basename(path)]
        seed = int(hashlib.md5(os.path.basename(path).encode()).hexdigest()[:8], 16)
        g = torch.Generator().manual_seed(seed)
        ch = 2 if os.path.basename(path) in stereo else 1
        return torch.randn((ch, n), generator=g) * 0.1, 16000
    return load


def build(root):
    """Create the files under `root` (a pathlib.Path); returns (speaker_dir, noise_json, loader)."""
    rng = np.random.default_rng(0)
    spk = root / "spk" / "chapter"
    spk.mkdir(parents=True)
    lengths = {}
    for i in range(14):                                   # utterances of 2-15 s plus a transcript file
        name = "61-%04d.flac" % i
        (spk / name).write_bytes(b"")
        lengths[name] = int(rng.integers(32000, 240000))
    (spk / "61.trans.txt").write_text("x")
    bg = {}
    for i in range(6):
        name = str(root / ("noise_%d.wav" % i))
        bg[name] = int(rng.integers(100000, 700000))
        lengths[os.path.basename(name)] = bg[name]
    (root / "noise.json").write_text(json.dumps(bg))
    return str(root / "spk"), str(root / "noise.json"), fake_loader(lengths, stereo=("noise_1.wav", "noise_4.wav"))


def digest(t):
    """sha256 of a float32 tensor's bytes."""
    return hashlib.sha256(t.detach().cpu().contiguous().numpy().astype(np.float32).tobytes()).hexdigest()


def sorted_walk(real_walk):
    """Directory order is filesystem business: give every implementation the same sorted listing."""
    return lambda top, *a, **k: [(r, sorted(d), sorted(f)) for r, d, f in real_walk(top, *a, **k)]
