"""Dry-stream assembly: drop-ins for SonicSim_audio.get_random_wav_path (:152-190),
get_random_wav_path_from_json (:192-228), create_long_audio (:231-279) and create_background_audio (:281-340).

Host logic only (file selection with the `random` module, placement with random gaps, stereo -> mono,
resampling): there is no arithmetic to accelerate here; the functions exist so that SonicSet.py:72-83 finds
every name it calls on the drop-in module.  They consume the `random` module's stream in exactly the
reference's order, so a seeded run selects and places the same files.

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


def create_long_audio(audio_path: str, length: float, sample_rate: int = 16000, loader=None):
    """:231-279 - utterances of one speaker laid end to end, each preceded by 0-10 s of silence.
    Returns (long_audio (1, length*sr), [(start, end), ...], [path, ...])."""
    import torch
    loader = loader or _default_loader
    print("create_long_audio: ", audio_path)
    total = int(length * sample_rate)
    paths = get_random_wav_path(audio_path, total, loader=loader)
    clips = [_load_resampled(p, sample_rate, loader) for p in paths]
    long_audio = torch.zeros((1, total), device=clips[0].device)
    spans, used, cursor = [], [], 0
    while cursor < total and clips:
        k = random.randint(0, len(clips) - 1)
        gap = random.randint(0, int(10 * sample_rate))
        n = gap + clips[k].shape[-1]
        if cursor + n > total:
            break
        spans.append((cursor + gap, cursor + n))
        long_audio[:, cursor + gap:cursor + n] += clips[k]
        cursor += n
        used.append(paths.pop(k))
        clips.pop(k)
    return long_audio, spans, used


def create_background_audio(audio_path: str, length: float, sample_rate: int = 16000, loader=None):
    """:281-340 - noise / music bed from a {path: n_samples} JSON: clips (stereo averaged to mono) followed by
    0-10 s of silence; the clip that reaches the end is trimmed by up to 10 % of the remainder on both sides."""
    import torch
    loader = loader or _default_loader
    print("create_background_audio: ", audio_path)
    total = int(length * sample_rate)
    paths = get_random_wav_path_from_json(audio_path, total, threshold=0.4)
    clips = [_load_resampled(p, sample_rate, loader) for p in paths]
    long_audio = torch.zeros((1, total), device=clips[0].device)
    spans, used, cursor = [], [], 0
    while cursor < total and clips:
        k = random.randint(0, len(clips) - 1)
        clip = clips[k]
        if clip.shape[0] == 2:
            clip = clip.mean(dim=0, keepdim=True)
        tail = random.randint(0, int(10 * sample_rate))
        n = clip.shape[-1] + tail
        if n >= total - cursor:                                   # this clip reaches the end of the bed
            slack = int((length * sample_rate - cursor) * 0.1)
            lead, trail = random.randint(0, slack), random.randint(0, slack)
            spans.append((lead + cursor, total - trail))
            used.append(paths.pop(k))
            clips.pop(k)
            padded = torch.cat([clip, torch.zeros((1, tail), device=clip.device)], dim=-1)
            try:
                long_audio[:, lead + cursor:total - trail] += padded[:, lead:total - trail - cursor]
            except Exception:                                     # noqa: BLE001  (the reference's bare except, :324-328)
                pass
            break
        spans.append((cursor, cursor + n))
        long_audio[:, cursor:cursor + clip.shape[-1]] += clip
        cursor += n
        used.append(paths.pop(k))
        clips.pop(k)
    return long_audio, spans, used
