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


def _np_dtype(a):
    """dtype NumPy would see for `a` (torch tensors included), for the output-dtype rule of the reference's NumPy code."""
    if hasattr(a, "detach"):
        return np.dtype(str(a.dtype).replace("torch.", ""))
    return np.asarray(a).dtype


def _out_dtype(*arrays):
    """The reference runs scipy / NumPy on whatever it is given, so float64 in means float64 out.  The CUDA path
    computes in float32 (the SonicSet pipeline is float32 end to end); a float64 caller gets the float32
    result widened to float64 - same dtype as the reference, float32 accuracy."""
    dt = np.result_type(np.float32, *[d for d in map(_np_dtype, arrays) if d.kind in "fc"])
    return np.dtype(np.float64) if dt == np.float64 else np.dtype(np.float32)


def convolve_fixed_receiver(source_audio, rirs) -> np.ndarray:
    """SonicSim_moving.py:47-61: `fftconvolve(x.reshape(1,-1), rirs, 'full')[:, :N]` -> (C, N)."""
    out_dtype = _out_dtype(source_audio, rirs)
    x = _as_f32(source_audio).reshape(-1)
    h = _as_f32(rirs)
    if h.ndim != 2:
        raise ValueError("rirs must have shape (num_channels, ir_length)")
    C, L = h.shape
    N = x.shape[0]
    out = np.empty((C, N), dtype=np.float32)
    if N == 0:
        return out.astype(out_dtype, copy=False)
    it = SsSource(x=x.ctypes.data, rir=h.ctypes.data, out=out.ctypes.data, N=N, P=1, C=C, L=L, mode=_lib.SS_STATIC)
    _render_one(it, (x, h, out))
    return out.astype(out_dtype, copy=False)


def convolve_moving_receiver(source_audio: np.ndarray, rirs: np.ndarray, interp_index, interp_weight) -> np.ndarray:
    """SonicSim_moving.py:63-96: (1 - w) * conv[idx] + w * conv[idx + 1] per sample -> (C, N)."""
    out_dtype = _out_dtype(source_audio, rirs, interp_weight)
    x = _as_f32(source_audio)
    h = _as_f32(rirs)
    if x.ndim != 1 or h.ndim != 3:
        raise ValueError("source_audio must be (audio_len,), rirs (num_positions, num_channels, ir_length)")
    P, C, L = h.shape
    N = x.shape[0]
    idx = np.asarray(interp_index)
    if idx.dtype.kind not in "iu":                 # NumPy's own rule for the fancy index at :89
        raise IndexError("arrays used as indices must be of integer (or boolean) type")
    w = np.ascontiguousarray(interp_weight, dtype=np.float32)
    if idx.shape != (N,) or w.shape != (N,):
        raise IndexError("interp_index / interp_weight must have shape (audio_len,)")
    out = np.empty((C, N), dtype=np.float32)
    if N == 0:
        return out.astype(out_dtype, copy=False)
    # NumPy fancy indexing of :89-90: `conv[idx]` and `conv[idx + 1]` each accept [-P, P); negative values wrap.
    lo, hi = int(idx.min()), int(idx.max())
    if lo < -P or hi + 1 >= P:
        bad = lo if lo < -P else hi + 1
        raise IndexError("index %d is out of bounds for axis 0 with size %d" % (bad, P))
    if lo < 0:
        # idx in [-P, -2] names the pair (idx + P, idx + P + 1); idx == -1 the pair (P - 1, 0): give the kernel a
        # (P + 1)-th position that repeats position 0 so that every pair is (p, p + 1) again
        idx = idx.astype(np.int64)
        if (idx == -1).any():
            h = np.ascontiguousarray(np.concatenate([h, h[:1]], axis=0))
            P += 1
            idx = np.where(idx < 0, idx + (P - 1), idx)
        else:
            idx = np.where(idx < 0, idx + P, idx)
    idx32 = np.ascontiguousarray(idx, dtype=np.int32)
    it = SsSource(x=x.ctypes.data, rir=h.ctypes.data, out=out.ctypes.data, idx=idx32.ctypes.data,
                  w=w.ctypes.data, N=N, P=P, C=C, L=L, mode=_lib.SS_MOVING_INDEXED)
    _render_one(it, (x, h, out, idx32, w))
    return out.astype(out_dtype, copy=False)


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
    if bounds.shape[0] <= P:
        # more RIRs than waypoints: interp_index never reaches the extra ones (the reference just ignores them)
        P = bounds.shape[0]
        h = np.ascontiguousarray(h[:P])
    else:
        # fewer RIRs than waypoints: the reference raises IndexError as soon as a sample needs position >= P
        used = np.nonzero(np.diff(bounds))[0]
        if used.size and used[-1] + 1 >= P:
            raise IndexError("index %d is out of bounds for axis 0 with size %d" % (int(used[-1]) + 1, P))
        bounds = np.ascontiguousarray(bounds[:P])
    out = np.empty((C, audio_len), dtype=np.float32)
    it = SsSource(x=x.ctypes.data, rir=h.ctypes.data, out=out.ctypes.data, bounds=bounds.ctypes.data,
                  N=audio_len, P=P, C=C, L=L, mode=_lib.SS_MOVING_BOUNDS)
    _render_one(it, (x, h, out, bounds))
    return torch.from_numpy(out)


# helpers of the reference module that do not touch the GPU path (SonicSim_moving.py:127-144)
def interpolate_values(start: float, end: float, interp_weight: float) -> float:
    """SonicSim_moving.py:127-144."""
    return (1 - interp_weight) * start + interp_weight * end
