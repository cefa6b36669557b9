"""Batch-level entry points (additive - NOT reference API).

`convolve_moving` and `render_scene` are the names BASELINE.json's north star uses; the reference
has no such symbols.  They render many (utterance, source) units with one C-ABI call
(ss_render_host / ss_render_dev) instead of the reference's serial per-source loop
(SonicSet.py:77-94).  Host arrays in / out by default; `Renderer.render_device` takes CUDA tensors.
"""
import ctypes
import typing as T

import numpy as np

from . import _lib
from ._lib import SsPostLufs, SsSource
from .SonicSim_moving import _as_f32, _samples_per_interval, bounds_from_counts


class MovingSource(T.NamedTuple):
    """One moving source: dry (N,), RIRs (P, C, L), trajectory as int32 segment bounds (P,).
    `bounds_host` (device path only, optional): NumPy copy of `bounds`; lets the library build the block
    table on the host (exact choice of the blocking plan, one small kernel less).  It must hold the same values as
    the device tensor whenever the batch runs - a plan that is re-run after `bounds` was overwritten on the device
    needs a new plan (or no `bounds_host`)."""
    dry: T.Any
    rirs: T.Any
    bounds: T.Any
    bounds_host: T.Any = None
    normalize_rirs: bool = False      # rirs = raw simulator output (formats.stack_rirs): divide by their global
                                      # abs-max on the device like generate_rir_combination (SonicSim_audio.py:398)


class StaticSource(T.NamedTuple):
    """One static source: dry (N,), RIR (C, L)."""
    dry: T.Any
    rir: T.Any
    normalize_rirs: bool = False


class Renderer:
    """Owns the per-device ss_ctx.  One process per GPU (LOCAL_RANK picks the device by default)."""

    def __init__(self, device: T.Optional[int] = None):
        self.lib = _lib.load()
        self.device = _lib.default_device() if device is None else int(device)
        self.ctx = _lib.context(self.device)

    # ------------------------------------------------------------------ host buffers
    def plan_host(self, sources: T.Sequence[T.Union[MovingSource, StaticSource]],
                  outs: T.Optional[T.Sequence[np.ndarray]] = None,
                  lufs_targets: T.Optional[T.Sequence[T.Optional[float]]] = None, sr: float = 16000) -> "HostPlan":
        """Validate a batch once and bind it to its host buffers; `plan.run()` then renders it with a single
        C-ABI call (a generation loop refills the same pinned buffers and calls run() again).
        See render_host for the arguments."""
        from .SonicSim_audio import gating_plan
        n = len(sources)
        items = (SsSource * n)()
        post = (SsPostLufs * n)() if lufs_targets is not None else None
        results_l = np.full((n, 2), np.nan, dtype=np.float64)
        keep = []
        results = []
        for i, s in enumerate(sources):
            x = _as_f32(s.dry).reshape(-1)
            if isinstance(s, StaticSource):
                h = _as_f32(s.rir)
                C, L = h.shape
                P, mode, bounds = 1, _lib.SS_STATIC, None
            else:
                h = _as_f32(s.rirs)
                P, C, L = h.shape
                bounds = np.ascontiguousarray(s.bounds, dtype=np.int32)
                mode = _lib.SS_MOVING_BOUNDS
                if bounds.shape != (P,):
                    raise ValueError("bounds must have P entries (cumulative segment bounds with leading 0)")
            N = x.shape[0]
            out = outs[i] if outs is not None else np.empty((C, N), dtype=np.float32)
            if out.shape != (C, N) or out.dtype != np.float32 or not out.flags.c_contiguous:
                raise ValueError("out[%d] must be C-contiguous float32 of shape (C, N)" % i)
            items[i] = SsSource(x=x.ctypes.data, rir=h.ctypes.data, out=out.ctypes.data,
                                bounds=bounds.ctypes.data if bounds is not None else None,
                                N=N, P=P, C=C, L=L, mode=mode, flags=_lib.SS_RIR_NORMALIZE if s.normalize_rirs else 0)
            keep.append((x, h, bounds, out))
            results.append(out)
            if post is not None and lufs_targets[i] is not None:
                if C > 8:
                    raise ValueError("loudness normalisation supports at most 8 channels")
                block = 0.4 if N / sr >= 0.4 else N / sr                      # SonicSim_audio.py:69
                brk, blo, bhi = gating_plan(N, float(sr), float(block))
                post[i] = SsPostLufs(brk=brk.ctypes.data, blk_lo=blo.ctypes.data, blk_hi=bhi.ctypes.data,
                                     n_e=len(brk) - 1, n_blocks=len(blo), rate=float(sr), block_size=float(block),
                                     target_lufs=float(lufs_targets[i]), result=results_l[i].ctypes.data)
                keep.append((brk, blo, bhi))
        return HostPlan(self, items, post, n, results, results_l, keep)

    def render_host(self, sources: T.Sequence[T.Union[MovingSource, StaticSource]],
                    outs: T.Optional[T.Sequence[np.ndarray]] = None,
                    lufs_targets: T.Optional[T.Sequence[T.Optional[float]]] = None, sr: float = 16000,
                    loudness_out: T.Optional[list] = None) -> T.List[np.ndarray]:
        """Render every source; returns a list of (C, N) float32 arrays (pinned `outs` may be passed).
        `lufs_targets[i]` (or None) additionally normalises stem i to that integrated loudness on
        the device before the copy out - SonicSim_audio.lufs_norm fused behind the render
        (SonicSet.py:97-101).  `loudness_out`, if a list, receives (measured LUFS, linear gain) per source."""
        plan = self.plan_host(sources, outs, lufs_targets, sr)
        res = plan.run()
        if loudness_out is not None and lufs_targets is not None:
            loudness_out[:] = plan.loudness()
        return res

    # ------------------------------------------------------------------ device tensors
    def render_device(self, sources, outs, stream: T.Optional[int] = None):
        """`sources`: MovingSource / StaticSource whose fields are CUDA float32 / int32 tensors;
        `outs`: preallocated CUDA (C, N) tensors.  Asynchronous on `stream` (torch current stream)."""
        items, n, keep = self._device_items(sources, outs)
        if stream is None:
            import torch
            stream = torch.cuda.current_stream().cuda_stream
        _lib.check(self.lib.ss_render_dev(self.ctx, items, n, ctypes.c_void_p(stream)))

    def plan_device(self, sources, outs) -> "DevicePlan":
        """Bind a batch of device-resident sources to its output tensors once; `plan.run()` then renders it with a
        single C-ABI call - one CUDA graph launch (ss_plan_run; a generation loop overwrites the same tensors and
        calls run() again).  Every tensor must be a contiguous CUDA tensor of this renderer's device: float32 dry
        (N,), RIRs (P, C, L) / (C, L), outs[i] (C, N) for that source's own C and N; int32 bounds (P,)."""
        return DevicePlan(self, *self._device_items(sources, outs))

    def _device_items(self, sources, outs):
        """Validated ss_source array of a device-resident batch: (items, n, objects to keep alive)."""
        import torch
        n = len(sources)
        if len(outs) != n:
            raise ValueError("plan_device: %d sources but %d output tensors" % (n, len(outs)))
        items = (SsSource * n)()
        keep = [sources, outs]

        def chk(t, name, i, dtype, shape=None):
            if not isinstance(t, torch.Tensor) or not t.is_cuda:
                raise ValueError("plan_device: %s of source %d must be a CUDA tensor" % (name, i))
            if self.device is not None and t.device.index != self.device:
                raise ValueError("plan_device: %s of source %d is on cuda:%s, the renderer on cuda:%d"
                                 % (name, i, t.device.index, self.device))
            if t.dtype != dtype:
                raise ValueError("plan_device: %s of source %d must be %s, got %s" % (name, i, dtype, t.dtype))
            if not t.is_contiguous():
                raise ValueError("plan_device: %s of source %d must be contiguous" % (name, i))
            if shape is not None and tuple(t.shape) != tuple(shape):
                raise ValueError("plan_device: %s of source %d must have shape %s, got %s"
                                 % (name, i, tuple(shape), tuple(t.shape)))

        for i, s in enumerate(sources):
            chk(s.dry, "dry", i, torch.float32)
            N = s.dry.numel()
            if isinstance(s, StaticSource):
                chk(s.rir, "rir", i, torch.float32)
                if s.rir.dim() != 2:
                    raise ValueError("plan_device: rir of static source %d must be (C, L)" % i)
                C, L = s.rir.shape
                chk(outs[i], "out", i, torch.float32, (C, N))
                items[i] = SsSource(x=s.dry.data_ptr(), rir=s.rir.data_ptr(), out=outs[i].data_ptr(),
                                    N=N, P=1, C=C, L=L, mode=_lib.SS_STATIC,
                                    flags=_lib.SS_RIR_NORMALIZE if s.normalize_rirs else 0)
            else:
                chk(s.rirs, "rirs", i, torch.float32)
                if s.rirs.dim() != 3:
                    raise ValueError("plan_device: rirs of moving source %d must be (P, C, L)" % i)
                P, C, L = s.rirs.shape
                chk(outs[i], "out", i, torch.float32, (C, N))
                chk(s.bounds, "bounds", i, torch.int32, (P,))
                bh = None
                if s.bounds_host is not None:
                    bh = np.ascontiguousarray(s.bounds_host, dtype=np.int32)
                    if bh.shape != (P,):
                        raise ValueError("plan_device: bounds_host of source %d must have P entries" % i)
                    keep.append(bh)
                items[i] = SsSource(x=s.dry.data_ptr(), rir=s.rirs.data_ptr(), out=outs[i].data_ptr(),
                                    bounds=s.bounds.data_ptr(), N=N, P=P, C=C, L=L,
                                    mode=_lib.SS_MOVING_BOUNDS, bounds_host=bh.ctypes.data if bh is not None else None,
                                    flags=_lib.SS_RIR_NORMALIZE if s.normalize_rirs else 0)
        return items, n, keep

    def check_device_errors(self):
        """Waits for the device and raises what the reference would have raised for a trajectory the device path met
        but could not refuse up front: IndexError for an interp_index outside [0, P - 2] (NumPy's fancy index at
        SonicSim_moving.py:89-90), ValueError for a device-side bounds table that is not ascending from 0 to N."""
        bits = ctypes.c_uint32()
        _lib.check(self.lib.ss_device_errors(self.ctx, ctypes.byref(bits)))
        if bits.value & 1:
            raise IndexError("an interp_index on the device was out of bounds for the number of RIR positions")
        if bits.value & 2:
            raise ValueError("a device-side trajectory bounds table is not ascending from 0 to N")

    def launch_count(self) -> int:
        return int(self.lib.ss_launch_count(self.ctx))

    def reset_stats(self):
        self.lib.ss_reset_stats(self.ctx)

    def set_profiling(self, on: bool):
        _lib.check(self.lib.ss_set_profiling(self.ctx, 1 if on else 0))

    def get_profile(self):
        """(ms in k_spectra, ms in k_render, launch pairs) since the previous call."""
        a, b, n = ctypes.c_double(), ctypes.c_double(), ctypes.c_int64()
        _lib.check(self.lib.ss_get_profile(self.ctx, ctypes.byref(a), ctypes.byref(b), ctypes.byref(n)))
        return a.value, b.value, n.value

    def set_chunk_bytes(self, nbytes: int):
        _lib.check(self.lib.ss_set_chunk_bytes(self.ctx, int(nbytes)))


class DevicePlan:
    """A batch of device-resident sources bound to its output tensors (Renderer.plan_device): an ss_plan of the C ABI -
    descriptors, block tables and scratch resident on the device, the launches captured into a CUDA graph."""

    def __init__(self, renderer, items, n, keep):
        self.renderer, self.items, self.n, self._keep = renderer, items, n, keep
        self._plan = ctypes.c_void_p()
        if n > 0:
            _lib.check(renderer.lib.ss_plan_create(renderer.ctx, items, n, ctypes.byref(self._plan)))

    def run(self, stream: T.Optional[int] = None):
        """Asynchronous on `stream` (default: torch's current stream)."""
        if self.n == 0:
            return
        if stream is None:
            import torch
            stream = torch.cuda.current_stream().cuda_stream
        _lib.check(self.renderer.lib.ss_plan_run(self._plan, ctypes.c_void_p(stream)))

    def is_graph(self) -> bool:
        return self.n > 0 and self.renderer.lib.ss_plan_is_graph(self._plan) == 1

    def close(self):
        if self._plan:
            self.renderer.lib.ss_plan_destroy(self._plan)
            self._plan = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class HostPlan:
    """A validated batch bound to its host input / output buffers (Renderer.plan_host)."""

    def __init__(self, renderer, items, post, n, results, results_l, keep):
        self.renderer, self.items, self.post, self.n = renderer, items, post, n
        self.results, self.results_l, self._keep = results, results_l, keep

    def run(self) -> T.List[np.ndarray]:
        """H2D -> render [-> loudness] -> D2H of the whole batch; returns the (C, N) output arrays."""
        lib, ctx = self.renderer.lib, self.renderer.ctx
        if self.post is None:
            _lib.check(lib.ss_render_host(ctx, self.items, self.n))
        else:
            _lib.check(lib.ss_render_host_ex(ctx, self.items, self.n, self.post))
        return self.results

    def loudness(self) -> T.List[T.Tuple[float, float]]:
        """(measured LUFS, linear gain) per source of the last run (NaN where no target was given)."""
        return [tuple(r) for r in self.results_l]


_default: T.Optional[Renderer] = None


def default_renderer() -> Renderer:
    global _default
    if _default is None:
        _default = Renderer()
    return _default


def trajectory_bounds(receiver_position, total_samples: int) -> np.ndarray:
    """Positions (P, 3) -> int32 segment bounds (P,), via the reference's constant-speed rule
    (SonicSim_moving.py:32-39; consumes the global NumPy RNG like the reference)."""
    return bounds_from_counts(_samples_per_interval(np.asarray(receiver_position), total_samples))


def convolve_moving(dry_list, rirs_list, positions_list) -> T.List[np.ndarray]:
    """Batched interpolate_moving_audio: for each i, dry (N,), RIRs (P, C, L), positions (P, 3)."""
    srcs = [MovingSource(d, r, trajectory_bounds(p, np.asarray(d).shape[-1]))
            for d, r, p in zip(dry_list, rirs_list, positions_list)]
    return default_renderer().render_host(srcs)


def render_scene(moving: T.Sequence[T.Tuple], static: T.Sequence[T.Tuple] = (), sr: float = 16000,
                 moving_lufs: T.Optional[float] = None, static_lufs: T.Optional[T.Sequence[float]] = None
                 ) -> T.Tuple[T.List[np.ndarray], T.List[np.ndarray]]:
    """One scene of SonicSet.process_single (SonicSet.py:77-101) in a single call:
    `moving` = [(dry (N,), rirs (P, C, L), positions (P, 3)), ...], `static` = [(dry (N,), rir (C, L)), ...].
    With `moving_lufs` (SonicSet uses -17) / `static_lufs` (SonicSet: [-24, -29] for noise, music) every
    stem is loudness-normalised like get_lufs_norm_audio: target ~ U(lufs - 2, lufs + 2) from the global
    NumPy RNG, drawn in the reference's order (all moving stems, then the static ones).
    Returns (moving stems, static stems), each (C, N) float32."""
    srcs: T.List = [MovingSource(d, r, trajectory_bounds(p, np.asarray(d).shape[-1])) for d, r, p in moving]
    srcs += [StaticSource(d, r) for d, r in static]
    targets = None
    if moving_lufs is not None or static_lufs is not None:
        targets = []
        for _ in moving:
            targets.append(None if moving_lufs is None else np.random.uniform(moving_lufs - 2, moving_lufs + 2))
        for i, _ in enumerate(static):
            l = None if static_lufs is None else static_lufs[i]
            targets.append(None if l is None else np.random.uniform(l - 2, l + 2))
    outs = default_renderer().render_host(srcs, lufs_targets=targets, sr=sr)
    return outs[: len(moving)], outs[len(moving):]


def render_mixtures(utterances: T.Sequence[dict], device: T.Optional[int] = None) -> T.List[dict]:
    """Whole path on the device, per utterance: render every speaker / noise stem, then assemble the training
    mixture exactly as the dataloader does (separation/look2hear/datas/movingdatamodule.py:105-124) without the
    stems leaving HBM.  Additive API (torch is used for device memory and copies only).

    utterance = {"speakers": [(dry (N,), rirs (P, C, L), positions (P, 3)), ...],     # speaker 0 is the reference
                 "noises":   [(dry (N,), rir (C, L)), ...],
                 "sirs": (S-1,) dB, "snr": dB}
    Returns per utterance {"mix": (C, N) float32, "speakers": (S, C, N) float32 (after their gains)}.
    Every stem of an utterance must have the same (C, N)."""
    import torch
    from ._lib import SsMixItem
    R = default_renderer() if device is None else Renderer(device)
    dev = torch.device("cuda", R.device if R.device is not None else torch.cuda.current_device())
    lib = R.lib
    srcs, outs, plan = [], [], []
    for u in utterances:
        S, M = len(u["speakers"]), len(u["noises"])
        N = int(np.asarray(u["speakers"][0][0]).shape[-1])
        C = int(np.asarray(u["speakers"][0][1]).shape[1])
        for d, h, *_ in list(u["speakers"]) + list(u["noises"]):
            hs = np.asarray(h).shape
            if int(np.asarray(d).shape[-1]) != N or int(hs[-2]) != C:
                raise ValueError("render_mixtures: every stem of an utterance must have the same (C, N); got N=%d C=%d "
                                 "against N=%d C=%d" % (int(np.asarray(d).shape[-1]), int(hs[-2]), N, C))
        spk = torch.empty((S, C, N), dtype=torch.float32, device=dev)
        noi = torch.empty((M, C, N), dtype=torch.float32, device=dev)
        for i, (d, h, pos) in enumerate(u["speakers"]):
            b = trajectory_bounds(pos, N)
            srcs.append(MovingSource(torch.from_numpy(_as_f32(d).reshape(-1)).to(dev), torch.from_numpy(_as_f32(h)).to(dev),
                                     torch.from_numpy(b).to(dev), b))
            outs.append(spk[i])
        for i, (d, h) in enumerate(u["noises"]):
            srcs.append(StaticSource(torch.from_numpy(_as_f32(d).reshape(-1)).to(dev), torch.from_numpy(_as_f32(h)).to(dev)))
            outs.append(noi[i])
        plan.append((spk, noi, S, M, C, N, u))
    R.render_device(srcs, outs)
    n = len(plan)
    items = (SsMixItem * n)()
    keep = []
    nscr = int(lib.ss_mix_scratch_doubles())
    for k, (spk, noi, S, M, C, N, u) in enumerate(plan):
        mix = torch.empty((C, N), dtype=torch.float32, device=dev)
        sirs = torch.tensor(np.asarray(u.get("sirs", np.zeros(max(S - 1, 1))), dtype=np.float32).reshape(-1), device=dev)
        scr = torch.empty(nscr, dtype=torch.float64, device=dev)
        items[k] = SsMixItem(speakers=spk.data_ptr(), noises=noi.data_ptr(), sirs=sirs.data_ptr(), mix=mix.data_ptr(),
                             speakers_out=spk.data_ptr(), scratch=scr.data_ptr(), E=C * N, S=S, M=M,
                             snr=float(np.asarray(u.get("snr", 15.0)).reshape(-1)[0]))
        keep.append((mix, sirs, scr))
    _lib.check(lib.ss_mix_dev(R.ctx, items, n, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
    torch.cuda.synchronize()
    return [{"mix": keep[k][0].cpu().numpy(), "speakers": plan[k][0].cpu().numpy()} for k in range(n)]
