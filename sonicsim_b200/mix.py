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


# ---- enhancement dataloader variant (enhancement/look2hear/datas/movingdatamodule.py) -------------------------
def overlap_audio(waveform, sample_rate, delay=6):
    """:34-48, same signature: waveform (rows, T) float32 tensor -> waveform + itself shifted by +-delay seconds
    (zeros shifted in), summed in the reference's order.  Runs on the GPU (`ss_overlap_host`)."""
    import torch
    if waveform.dim() != 2:
        raise IndexError("Dimension out of range (expected to be in range of [-2, 1], but got 1)")   # .size(1) at :42
    x = np.ascontiguousarray(waveform.detach().cpu().numpy(), dtype=np.float32)
    y = np.empty_like(x)
    delay_samples = int(delay * sample_rate)                                                      # :35
    if delay_samples < 0:
        raise ValueError("negative delay")
    _lib.check(_lib.load().ss_overlap_host(_lib.context(), x.ctypes.data, y.ctypes.data, x.shape[0], x.shape[1],
                                           delay_samples))
    return torch.from_numpy(y)


def find_overlap_region(data, min_overlap=2, max_overlap=3, max_duration=None, sample_rate=None):
    """:50-75, same signature and the same `random.randint` draws (host logic): a region [start, end] of the scene
    holding between `min_overlap` and `max_overlap` utterance boundaries of json_data.json's `start_end_points`.
    As in the reference, `max_duration` acts as a minimum length in seconds."""
    import random
    points = [pt for source in data.values() if "start_end_points" in source for pt in source["start_end_points"]]
    if not points:
        raise ValueError("min() arg is an empty sequence")                                       # :57
    first = min(pt[0] for pt in points)
    last = max(pt[1] for pt in points)
    check_len = max_duration is not None and sample_rate is not None
    while True:
        start = random.randint(first, last)
        end = random.randint(start, last)
        if check_len and (end - start) / sample_rate < max_duration:
            continue
        inside = 0
        for pt in points:
            inside += (start <= pt[0] <= end) or (start <= pt[1] <= end)
        if min_overlap <= inside <= max_overlap:
            return start, end


def mix_noisy(speaker_wavs, noise_wav, snr: T.Optional[float] = None, sample_rate=16000, delay=6):
    """The noisy single-speaker mixture of the enhancement dataloader (:235-257): sum of the noise stems,
    `overlap_audio(delay)` over the flattened sum, SNR gain against the speaker stem clamped at +40 dB, sum.
    speaker_wavs (..., T), noise_wav (M, ..., T) float32 tensors -> mix_wav.  `snr=None` draws
    `torch.Tensor(1).uniform_(-10, 15)` like :244 / :250.  One fused GPU pass set (`ss_mix_host_ex`)."""
    import torch
    if snr is None:
        snr = torch.Tensor(1).uniform_(-10, 15).numpy()
    spk = np.ascontiguousarray(speaker_wavs.detach().cpu().numpy(), dtype=np.float32)
    noi = np.ascontiguousarray(noise_wav.detach().cpu().numpy(), dtype=np.float32)
    if noi.ndim < 1 or spk.shape != noi.shape[1:]:
        raise RuntimeError("The size of tensor a must match the size of tensor b")               # broadcast at :257
    E = int(spk.size)
    delay_samples = int(delay * sample_rate)
    if delay_samples < 0:
        raise ValueError("negative delay")
    mix = np.empty(spk.shape, dtype=np.float32)
    st = _lib.load().ss_mix_host_ex(_lib.context(), spk.ctypes.data, noi.ctypes.data, None,
                                    float(np.asarray(snr).reshape(-1)[0]), mix.ctypes.data, None, 1, noi.shape[0], E,
                                    min(delay_samples, 2**31 - 1))
    _lib.check(st)
    return torch.from_numpy(mix)
