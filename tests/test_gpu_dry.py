"""SURVEY 8f row 2 on the device: create_long_audio / create_background_audio with `device=` (resampling, stereo -> mono
and placement in one CUDA kernel) against the host functions, which are pinned to the unmodified reference by
tests/test_dry.py."""
import hashlib
import json
import os
import random

import numpy as np
import pytest
import torch

from oracle import dry_fixture
from oracle import sonicsim_oracle as so
from sonicsim_b200 import dry

pytestmark = pytest.mark.gpu


def mixed_rate_loader(lengths):
    """Stand-in for torchaudio.load: rate and channel count are functions of the file name (44.1 kHz stereo, 48 kHz mono,
    16 kHz mono), content of its hash."""
    def load(path):
        name = os.path.basename(path)
        h = int(hashlib.md5(name.encode()).hexdigest()[:8], 16)
        i = int(name.split("_")[1].split(".")[0])
        sr, ch = ((44100, 2), (16000, 1), (48000, 1), (22050, 2), (44100, 1), (16000, 2))[i % 6]
        n = int(lengths[name] * sr / 16000)
        g = torch.Generator().manual_seed(h)
        return torch.randn((ch, n), generator=g) * 0.1, sr
    return load


def test_dry_streams_assembled_on_the_device(tmp_path, monkeypatch):
    spk, noise_json, load16 = dry_fixture.build(tmp_path)
    monkeypatch.setattr(os, "walk", dry_fixture.sorted_walk(os.walk))
    # 1. no resampling, mono: the device stream is bit-identical to the host (= reference) stream
    for seed in range(3):
        random.seed(seed)
        a, se, names = dry.create_long_audio(spk, 60, loader=load16)
        random.seed(seed)
        ad, sed, namesd = dry.create_long_audio(spk, 60, loader=load16, device="cuda")
        assert ad.is_cuda and ad.shape == (1, 960000) and se == sed and names == namesd
        assert torch.equal(ad.cpu(), a)
        random.seed(100 + seed)
        b, bse, bnames = dry.create_background_audio(noise_json, 60, loader=load16)        # two stereo clips in the fixture
        random.seed(100 + seed)
        bd, bsed, bnamesd = dry.create_background_audio(noise_json, 60, loader=load16, device="cuda")
        assert bse == bsed and bnames == bnamesd and torch.equal(bd.cpu(), b)
    # 2. mixed sample rates and channel counts: Resample's filter bank on the device, fp32 rounding apart
    lengths = {os.path.basename(k): v for k, v in json.load(open(noise_json)).items()}
    load = mixed_rate_loader(lengths)
    n_resampled = 0
    for seed in range(8):
        random.seed(200 + seed)
        b, bse, bnames = dry.create_background_audio(noise_json, 60, loader=load)
        random.seed(200 + seed)
        bd, bsed, bnamesd = dry.create_background_audio(noise_json, 60, loader=load, device="cuda")
        assert bse == bsed and bnames == bnamesd and bd.shape == b.shape
        n_resampled += sum(load(p)[1] != 16000 for p in bnames)
        err = so.rel_rms(bd.cpu().numpy(), b.numpy())
        assert err < 1e-6, err
        assert torch.equal(bd.cpu() == 0, b == 0) or err < 1e-6
    assert n_resampled > 0
    # 3. the device stream feeds the renderer's device path directly
    from sonicsim_b200 import render
    rng = np.random.default_rng(1)
    h = so.synth_rirs(rng, 1, 2, 900)[0]
    out = torch.empty((2, bd.shape[1]), device="cuda")
    render.default_renderer().render_device([render.StaticSource(bd[0], torch.from_numpy(h).cuda())], [out])
    torch.cuda.synchronize()
    assert so.rel_rms(out.cpu().numpy(), so.convolve_fixed_receiver(b.numpy(), h)) < 1e-4
