"""Dry-stream assembly: drop-ins for SonicSim_audio.get_random_wav_path (:152-190),
get_random_wav_path_from_json (:192-228), create_long_audio (:231-279) and create_background_audio (:281-340).

File selection and placement are host logic drawn from the `random` module in exactly the reference's order, so a
seeded run selects and places the same files.  The arithmetic - resampling, stereo -> mono, adding the clips into
the stream - runs on the CPU through torch like the reference by default, or, with `device=`, in one CUDA kernel
(ss_dry_assemble_dev, csrc/ss_dry.cu) that leaves the stream in HBM for the renderer's device path.

`loader(path) -> (Tensor (channels, n), sample_rate)` defaults to torchaudio.load.
"""
import json
import os
import random
import typing as T


def _default_loader(path):
    import torchaudio
    return torchaudio.load(path)


def _load_resampled(path, sample_rate, loader):
    import torchaudio
    waveform, sr = loader(path)
    if sr != sample_rate:
        waveform = torchaudio.transforms.Resample(orig_freq=sr, new_freq=sample_rate)(waveform)
    return waveform


def get_random_wav_path(audio_dir: str, length: int, threshold: float = 0.9, loader=None) -> T.List[str]:
    """:152-190 - random files of a directory tree whose total length lands in [threshold*length, length]."""
    loader = loader or _default_loader
    candidates = [os.path.join(root, name) for root, _, names in os.walk(str(audio_dir)) for name in names
                  if not name.endswith(".txt")]
    print(f"audio_path_list: {len(candidates)}")
    n_samples = {p: loader(p)[0].shape[-1] for p in candidates}
    chosen, total = [], 0
    while candidates and total < length * threshold:
        pick = random.choice(candidates)
        if total + n_samples[pick] > length:
            break
        chosen.append(pick)
        total += n_samples[pick]
        candidates.remove(pick)
    return chosen


def get_random_wav_path_from_json(json_dir: str, length: int, threshold: float = 0.9) -> T.List[str]:
    """:192-228 - same from a {path: n_samples} JSON; the file that overshoots is still taken, then it stops."""
    with open(json_dir) as f:
        n_samples = json.load(f)
    candidates = list(n_samples.keys())
    chosen, total = [], 0
    while candidates and total < length * threshold:
        pick = random.choice(candidates)
        chosen.append(pick)
        if not (total + n_samples[pick] < length):
            break
        total += n_samples[pick]
        candidates.remove(pick)
    return chosen


def _resampled_len(n: int, sr: int, sample_rate: int) -> int:
    """Length of torchaudio.transforms.Resample(sr, sample_rate)(x) for x of n samples: ceil(new * n / orig)."""
    if sr == sample_rate:
        return n
    import math
    g = math.gcd(int(sr), int(sample_rate))
    o, nw = int(sr) // g, int(sample_rate) // g
    return -((-nw * n) // o)


def _place_long_audio(lens: T.List[int], total: int, sample_rate: int):
    """The placement loop of create_long_audio (:258-277) on clip lengths alone; same `random` draws.
    Returns (placements [(clip, dst_start, src_start, count)], spans, order of clips used)."""
    alive = list(range(len(lens)))
    places, spans, used, cursor = [], [], [], 0
    while cursor < total and alive:
        k = random.randint(0, len(alive) - 1)
        gap = random.randint(0, int(10 * sample_rate))
        n = gap + lens[alive[k]]
        if cursor + n > total:
            break
        spans.append((cursor + gap, cursor + n))
        places.append((alive[k], cursor + gap, 0, lens[alive[k]]))
        cursor += n
        used.append(alive.pop(k))
    return places, spans, used


def _place_background_audio(lens: T.List[int], total: int, length: float, sample_rate: int):
    """The placement loop of create_background_audio (:305-338) on clip lengths alone; same `random` draws."""
    alive = list(range(len(lens)))
    places, spans, used, cursor = [], [], [], 0
    while cursor < total and alive:
        k = random.randint(0, len(alive) - 1)
        ln = lens[alive[k]]
        tail = random.randint(0, int(10 * sample_rate))
        n = ln + tail
        if n >= total - cursor:                                   # this clip reaches the end of the bed
            slack = int((length * sample_rate - cursor) * 0.1)
            lead, trail = random.randint(0, slack), random.randint(0, slack)
            spans.append((lead + cursor, total - trail))
            # long_audio[:, lead+cursor : total-trail] += padded[:, lead : total-trail-cursor]; beyond the clip the
            # padding is zeros (the slices always have equal length here, so the reference's bare except never fires)
            count = min(ln, total - trail - cursor) - lead
            if count > 0:
                places.append((alive[k], lead + cursor, lead, count))
            used.append(alive.pop(k))
            break
        spans.append((cursor, cursor + n))
        places.append((alive[k], cursor, 0, ln))
        cursor += n
        used.append(alive.pop(k))
    return places, spans, used


_kernel_cache: T.Dict[T.Tuple[int, int, str], T.Tuple] = {}


def _resample_kernel(sr: int, sample_rate: int, device):
    """torchaudio's own polyphase filter bank for Resample(sr, sample_rate), transposed to (taps, new) on `device`."""
    key = (int(sr), int(sample_rate), str(device))
    if key not in _kernel_cache:
        import math
        import torchaudio
        rs = torchaudio.transforms.Resample(orig_freq=sr, new_freq=sample_rate)
        g = math.gcd(int(sr), int(sample_rate))
        kt = rs.kernel[:, 0, :].t().contiguous().to(device)       # (taps, new)
        _kernel_cache[key] = (kt, int(sr) // g, int(sample_rate) // g, int(rs.width))
    return _kernel_cache[key]


def _assemble_device(raw, places, total: int, sample_rate: int, device):
    """Resample + stereo mean + placement of the drawn clips in one kernel (ss_dry_assemble_dev); returns a
    (1, total) float32 tensor on `device`."""
    import ctypes
    import torch
    from . import _lib
    from .render import Renderer, default_renderer
    dev = torch.device(device)
    if dev.type != "cuda":
        raise ValueError("dry-stream assembly on a device needs a CUDA device, got %r" % (device,))
    index = dev.index if dev.index is not None else torch.cuda.current_device()
    dev = torch.device("cuda", index)
    R = default_renderer()
    if R.device != index:                                         # a context of the library on the stream's own GPU
        R = Renderer(index)
    out = torch.empty((1, total), dtype=torch.float32, device=dev)
    n = len(places)
    clips = (_lib.SsDryClip * max(n, 1))()
    keep = []
    for i, (k, dst, src, count) in enumerate(places):
        wav, sr = raw[k]
        w = wav.to(dtype=torch.float32).contiguous()
        w = (w.pin_memory() if not w.is_cuda else w).to(dev, non_blocking=True)
        keep.append(w)
        kt, o, nw, width, taps = None, 1, 1, 0, 1
        if sr != sample_rate:
            kt, o, nw, width = _resample_kernel(sr, sample_rate, dev)
            taps = 2 * width + o
        clips[i] = _lib.SsDryClip(src=w.data_ptr(), kernel_t=kt.data_ptr() if kt is not None else None, dst_start=dst,
                                  src_start=src, count=count, channels=w.shape[0], src_len=w.shape[-1], orig=o,
                                  new_rate=nw, width=width, taps=taps)
    _lib.check(R.lib.ss_dry_assemble_dev(R.ctx, clips, n, ctypes.c_void_p(out.data_ptr()), total,
                                         ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
    for w in keep:                                               # the kernel reads the clips asynchronously
        w.record_stream(torch.cuda.current_stream(dev))
    return out


def create_long_audio(audio_path: str, length: float, sample_rate: int = 16000, loader=None, device=None):
    """:231-279 - utterances of one speaker laid end to end, each preceded by 0-10 s of silence.
    Returns (long_audio (1, length*sr), [(start, end), ...], [path, ...]).  With `device` (a CUDA device) the clips
    are resampled and placed by the GPU and the stream is returned as a device tensor; selection and placement
    consume the `random` stream exactly as without it."""
    import torch
    loader = loader or _default_loader
    print("create_long_audio: ", audio_path)
    total = int(length * sample_rate)
    paths = get_random_wav_path(audio_path, total, loader=loader)
    if device is not None:
        raw = [loader(p) for p in paths]
        lens = [_resampled_len(w.shape[-1], sr, sample_rate) for w, sr in raw]
        places, spans, used = _place_long_audio(lens, total, sample_rate)
        return _assemble_device(raw, places, total, sample_rate, device), spans, [paths[k] for k in used]
    clips = [_load_resampled(p, sample_rate, loader) for p in paths]
    long_audio = torch.zeros((1, total), device=clips[0].device)
    places, spans, used = _place_long_audio([c.shape[-1] for c in clips], total, sample_rate)
    for k, dst, src, count in places:
        long_audio[:, dst:dst + count] += clips[k][:, src:src + count]
    return long_audio, spans, [paths[k] for k in used]


def create_background_audio(audio_path: str, length: float, sample_rate: int = 16000, loader=None, device=None):
    """:281-340 - noise / music bed from a {path: n_samples} JSON: clips (stereo averaged to mono) followed by
    0-10 s of silence; the clip that reaches the end is trimmed by up to 10 % of the remainder on both sides.
    `device`: as for create_long_audio."""
    import torch
    loader = loader or _default_loader
    print("create_background_audio: ", audio_path)
    total = int(length * sample_rate)
    paths = get_random_wav_path_from_json(audio_path, total, threshold=0.4)
    if device is not None:
        raw = [loader(p) for p in paths]
        lens = [_resampled_len(w.shape[-1], sr, sample_rate) for w, sr in raw]
        places, spans, used = _place_background_audio(lens, total, length, sample_rate)
        return _assemble_device(raw, places, total, sample_rate, device), spans, [paths[k] for k in used]
    clips = [_load_resampled(p, sample_rate, loader) for p in paths]
    long_audio = torch.zeros((1, total), device=clips[0].device)
    places, spans, used = _place_background_audio([c.shape[-1] for c in clips], total, length, sample_rate)
    for k, dst, src, count in places:
        clip = clips[k]
        if clip.shape[0] == 2:
            clip = clip.mean(dim=0, keepdim=True)
        long_audio[:, dst:dst + count] += clip[:, src:src + count]
    return long_audio, spans, [paths[k] for k in used]
