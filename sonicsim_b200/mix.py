"""Mixture assembly of the training dataloader on the GPU.

Replaces the arithmetic of `MovingTrainDataset.__getitem__`
(separation/look2hear/datas/movingdatamodule.py:105-124; the enhancement tree has the same block at
enhancement/look2hear/datas/movingdatamodule.py:149-165): RMS-dB of the reference speaker, SIR gains of
the interferers and SNR gain of the summed noise (both clamped at +40 dB), sums.  File loading and the
random crop stay with the caller.  No CPU fallback.
"""
import typing as T

import numpy as np

from . import _lib


def compute_mch_rms_dB(mch_wav, fs=16000, energy_thresh=-50):
    """movingdatamodule.py:29-32 (host helper, same signature)."""
    import torch
    mean_square = max(1e-20, torch.mean(mch_wav ** 2))
    return 10 * np.log10(mean_square)


def mix_stems(speaker_wav, noise_wav, sirs: T.Optional[T.Sequence[float]] = None, snr: T.Optional[float] = None):
    """speaker_wav (S, ..., T), noise_wav (M, ..., T) float32 tensors (CPU) -> (mix_wav, speaker_wav scaled),
    exactly the values :105-124 produce.  When `sirs` / `snr` are None they are drawn like the reference
    (`torch.Tensor(S-1).uniform_(-6, 6)`, then `torch.Tensor(1).uniform_(10, 20)`), so a seeded torch RNG
    reproduces the reference's mixture."""
    import torch
    S, M = speaker_wav.shape[0], noise_wav.shape[0]
    if sirs is None:
        sirs = torch.Tensor(S - 1).uniform_(-6, 6).numpy()                       # :106
    if snr is None:
        snr = torch.Tensor(1).uniform_(10, 20).numpy()                           # :119
    spk = np.ascontiguousarray(speaker_wav.detach().cpu().numpy(), dtype=np.float32)
    noi = np.ascontiguousarray(noise_wav.detach().cpu().numpy(), dtype=np.float32)
    if spk.shape[1:] != noi.shape[1:]:
        raise RuntimeError("The size of tensor a must match the size of tensor b")  # torch's broadcast error at :123
    E = int(np.prod(spk.shape[1:]))
    sirs_f = np.ascontiguousarray(np.asarray(sirs, dtype=np.float32).reshape(-1))
    if sirs_f.shape[0] < S - 1:
        raise IndexError("index %d is out of bounds for axis 0 with size %d" % (sirs_f.shape[0], sirs_f.shape[0]))
    mix = np.empty(spk.shape[1:], dtype=np.float32)
    spk_out = np.empty_like(spk)
    st = _lib.load().ss_mix_host(_lib.context(), spk.ctypes.data, noi.ctypes.data, sirs_f.ctypes.data,
                                 float(np.asarray(snr).reshape(-1)[0]), mix.ctypes.data, spk_out.ctypes.data, S, M, E)
    _lib.check(st)
    return torch.from_numpy(mix), torch.from_numpy(spk_out)
