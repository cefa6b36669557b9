"""Drop-in for the reference module SonicSim-SonicSet/SonicSim_moving.py (functions :15-125).

Same names, positional signatures, return types and exception types, so that
`import SonicSim_moving` in SonicSet.py:20 can resolve to this module.
The arithmetic runs in the sm_100a CUDA library through the C ABI (include/sonicsim_b200.h);
host code here only adapts arrays.  No CPU fallback.
"""
import ctypes
import typing as T

import numpy as np

from . import _lib
from ._lib import SsSource


def _samples_per_interval(receiver_position: np.ndarray, total_samples: int) -> np.ndarray:
    """SonicSim_moving.py:32-39 - constant-speed sample counts per segment.  Consumes the global
    NumPy RNG exactly like the reference (`np.random.choice(S, |err|)`, :38)."""
    receiver_position = np.asarray(receiver_position)
    distance = np.linalg.norm(np.diff(receiver_position, axis=0), axis=1)
    speed_per_sample = distance.sum() / total_samples
    samples_per_interval = np.round(distance / speed_per_sample).astype(int)
    error = total_samples - samples_per_interval.sum()
    for i in np.random.choice(len(samples_per_interval), abs(error)):
        samples_per_interval[i] += np.sign(error)
    return samples_per_interval


def setup_dynamic_interp(receiver_position: np.ndarray, total_samples: int) -> T.Tuple[np.ndarray, np.ndarray]:
    """SonicSim_moving.py:15-45.  Host-side trajectory setup (O(P) + materialising (idx, w) because
    the reference signature returns them).  The renderer itself only needs the cumulative segment
    bounds; interpolate_moving_audio() below never builds these arrays."""
    spi = _samples_per_interval(receiver_position, total_samples)
    interp_index = np.repeat(np.arange(len(spi)), spi)                 # raises ValueError on negatives, like :42
    interp_weight = np.concatenate([np.linspace(0, 1, num, endpoint=False) for num in spi])
    return interp_index, interp_weight.astype(np.float32)


def _as_f32(a):
    if hasattr(a, "detach"):                     # torch.Tensor (SonicSet.py:93 passes tensors)
        a = a.detach().cpu().numpy()
    return np.ascontiguousarray(a, dtype=np.float32)


def _render_one(item: SsSource, keepalive):
    st = _lib.load().ss_render_host(_lib.context(), ctypes.byref(item), 1)
    _lib.check(st)


def convolve_fixed_receiver(source_audio, rirs) -> np.ndarray:
    """SonicSim_moving.py:47-61: `fftconvolve(x.reshape(1,-1), rirs, 'full')[:, :N]` -> (C, N)."""
    in_dtype = np.result_type(np.asarray(source_audio).dtype if not hasattr(source_audio, "detach") else np.float32, np.float32)
    x = _as_f32(source_audio).reshape(-1)
    h = _as_f32(rirs)
    if h.ndim != 2:
        raise ValueError("rirs must have shape (num_channels, ir_length)")
    C, L = h.shape
    N = x.shape[0]
    out = np.empty((C, N), dtype=np.float32)
    if N == 0:
        return out
    it = SsSource(x=x.ctypes.data, rir=h.ctypes.data, out=out.ctypes.data, N=N, P=1, C=C, L=L, mode=_lib.SS_STATIC)
    _render_one(it, (x, h, out))
    return out if in_dtype == np.float32 else out.astype(in_dtype)


def convolve_moving_receiver(source_audio: np.ndarray, rirs: np.ndarray, interp_index, interp_weight) -> np.ndarray:
    """SonicSim_moving.py:63-96: (1 - w) * conv[idx] + w * conv[idx + 1] per sample -> (C, N)."""
    x = _as_f32(source_audio)
    h = _as_f32(rirs)
    if x.ndim != 1 or h.ndim != 3:
        raise ValueError("source_audio must be (audio_len,), rirs (num_positions, num_channels, ir_length)")
    P, C, L = h.shape
    N = x.shape[0]
    idx = np.asarray(interp_index)
    w = np.ascontiguousarray(interp_weight, dtype=np.float32)
    if idx.shape != (N,) or w.shape != (N,):
        raise IndexError("interp_index / interp_weight must have shape (audio_len,)")
    out = np.empty((C, N), dtype=np.float32)
    if N == 0:
        return out
    # numpy fancy indexing semantics of :89-90: negative indices wrap, anything else out of range raises
    if idx.min() < -P or idx.max() + 1 >= P or (idx.min() < 0 and (idx + 1).max() >= P):
        raise IndexError("index %d is out of bounds for axis 0 with size %d" % (int(idx.max()) + 1, P))
    if idx.min() < 0:
        raise IndexError("negative interp_index is not supported by the CUDA path")
    idx32 = np.ascontiguousarray(idx, dtype=np.int32)
    it = SsSource(x=x.ctypes.data, rir=h.ctypes.data, out=out.ctypes.data, idx=idx32.ctypes.data,
                  w=w.ctypes.data, N=N, P=P, C=C, L=L, mode=_lib.SS_MOVING_INDEXED)
    _render_one(it, (x, h, out, idx32, w))
    return out


def bounds_from_counts(samples_per_interval) -> np.ndarray:
    """Compact trajectory: int32 cumulative segment bounds with a leading 0 (length P)."""
    spi = np.asarray(samples_per_interval)
    if (spi < 0).any():
        raise ValueError("repeats may not contain negative values.")       # np.repeat's message at :42
    return np.concatenate([[0], np.cumsum(spi)]).astype(np.int32)


def interpolate_moving_audio(source1_audio, ir1_list, receiver_position):
    """SonicSim_moving.py:98-125: Tensor (1, N), Tensor/list (P, 1, C, L), P positions -> Tensor (C, N)."""
    import torch
    audio_len = source1_audio.shape[-1]
    spi = _samples_per_interval(np.array(receiver_position), audio_len)          # :119
    bounds = bounds_from_counts(spi)
    x = _as_f32(source1_audio)[0]
    if isinstance(ir1_list, (list, tuple)):
        ir1_list = np.array([_as_f32(t) for t in ir1_list])
    h = _as_f32(ir1_list).squeeze(1)                                             # :122
    h = np.ascontiguousarray(h)
    P, C, L = h.shape
    if bounds.shape[0] != P:
        raise IndexError("number of receiver positions (%d) != number of RIRs (%d)" % (bounds.shape[0], P))
    out = np.empty((C, audio_len), dtype=np.float32)
    it = SsSource(x=x.ctypes.data, rir=h.ctypes.data, out=out.ctypes.data, bounds=bounds.ctypes.data,
                  N=audio_len, P=P, C=C, L=L, mode=_lib.SS_MOVING_BOUNDS)
    _render_one(it, (x, h, out, bounds))
    return torch.from_numpy(out)


# helpers of the reference module that do not touch the GPU path (SonicSim_moving.py:127-144)
def interpolate_values(start: float, end: float, interp_weight: float) -> float:
    """SonicSim_moving.py:127-144."""
    return (1 - interp_weight) * start + interp_weight * end
