"""Data formats either side of the render path - host-side plumbing.

  in : `rir_save_{mode}_{channel_type}.pt` - a list of per-speaker tensors (P, 1, C, L) written by
       SonicSet.py:68; `combine_rirs` is the post-processing of generate_rir_combination
       (SonicSim_audio.py:391-398: clip to the shortest, stack, divide by the global abs-max).
  out: (C, N) float32 stems as IEEE-float WAV exactly as torchaudio.save writes them in the reference
       (SonicSet.py:102-106; 18-byte fmt chunk, `fact` chunk, as in files/61-908-7127/*.wav) and the
       `json_data.json` sidecar (SonicSet.py:108-136).  `SceneWriter` overlaps the disk writes with
       rendering on a background thread.
"""
import json
import os
import queue
import struct
import threading
import typing as T

import numpy as np


# ------------------------------------------------------------------------------------ RIRs in
def clip_all(audio_list):
    """SonicSim_rir.py:24-41: clip every array / tensor to the shortest last dimension."""
    min_length = min(a.shape[-1] for a in audio_list)
    return [a[..., :min_length] for a in audio_list]


def combine_rirs(ir_list, num_sources: int, num_receivers: int = 1) -> np.ndarray:
    """SonicSim_audio.py:391-398: list of (C, L_i) impulse responses (one per source-receiver pair) ->
    float32 array (num_sources, num_receivers, C, L_min) normalised by its global abs-max."""
    irs = [np.asarray(a.detach().cpu().numpy() if hasattr(a, "detach") else a, dtype=np.float32) for a in ir_list]
    irs = clip_all(irs)
    num_channel = len(irs[0])
    out = np.stack(irs).reshape(num_sources, num_receivers, num_channel, -1)
    out = out / np.abs(out).max()
    return out.astype(np.float32)


def stack_rirs(ir_list, num_sources: int, num_receivers: int = 1) -> np.ndarray:
    """The clip + stack half of SonicSim_audio.py:391-397 WITHOUT the normalisation: pass the result to the renderer
    with `normalize_rirs=True` (ss_source.flags = SS_RIR_NORMALIZE) and the division by the global abs-max (:398) is
    done on the device while the RIRs are transformed (k_rir_absmax + k_prepare) - same bits as combine_rirs."""
    irs = [np.asarray(a.detach().cpu().numpy() if hasattr(a, "detach") else a, dtype=np.float32) for a in ir_list]
    irs = clip_all(irs)
    return np.ascontiguousarray(np.stack(irs).reshape(num_sources, num_receivers, len(irs[0]), -1))


def load_rir_dump(path: str) -> T.List[np.ndarray]:
    """`torch.save(ir_outputs, ...)` of SonicSet.py:68 -> list of (P, C, L) float32 arrays (receiver axis squeezed,
    as interpolate_moving_audio does at SonicSim_moving.py:122)."""
    import torch
    dump = torch.load(path, map_location="cpu")
    return [np.ascontiguousarray(t.numpy().squeeze(1), dtype=np.float32) for t in dump]


# ------------------------------------------------------------------------------------ WAV out
def wav_f32_bytes(stem: np.ndarray, sample_rate: int) -> bytes:
    """(C, N) float32 -> the bytes torchaudio.save(path, tensor (C, N), sample_rate) writes in the reference
    setup: WAVE_FORMAT_IEEE_FLOAT, 18-byte fmt chunk (cbSize 0), `fact` chunk with the frame count."""
    stem = np.asarray(stem, dtype=np.float32)
    if stem.ndim == 1:
        stem = stem[None, :]
    C, N = stem.shape
    data = np.ascontiguousarray(stem.T).tobytes()                     # interleaved frames
    fmt = struct.pack("<HHIIHHH", 3, C, sample_rate, sample_rate * C * 4, C * 4, 32, 0)
    fact = struct.pack("<I", N)
    body = b"WAVE" + b"fmt " + struct.pack("<I", len(fmt)) + fmt + b"fact" + struct.pack("<I", 4) + fact + \
           b"data" + struct.pack("<I", len(data)) + data
    return b"RIFF" + struct.pack("<I", len(body)) + body


def write_wav_f32(path: str, stem: np.ndarray, sample_rate: int) -> None:
    with open(path, "wb") as f:
        f.write(wav_f32_bytes(stem, sample_rate))


def read_wav_f32(path: str) -> T.Tuple[np.ndarray, int]:
    """Inverse of write_wav_f32 (float32 WAV only): returns ((C, N) float32, sample_rate)."""
    b = open(path, "rb").read()
    if b[:4] != b"RIFF" or b[8:12] != b"WAVE":
        raise ValueError("not a RIFF/WAVE file")
    pos, fmt, data = 12, None, None
    while pos + 8 <= len(b):
        cid, size = b[pos:pos + 4], struct.unpack("<I", b[pos + 4:pos + 8])[0]
        if cid == b"fmt ":
            fmt = struct.unpack("<HHIIHH", b[pos + 8:pos + 24])
        elif cid == b"data":
            data = b[pos + 8:pos + 8 + size]
        pos += 8 + size + (size & 1)
    if fmt is None or data is None or fmt[0] != 3 or fmt[5] != 32:
        raise ValueError("only IEEE float32 WAV is supported")
    x = np.frombuffer(data, dtype="<f4").reshape(-1, fmt[1]).T
    return np.ascontiguousarray(x), fmt[2]


def scene_json(sources: T.Sequence[dict], noise: T.Optional[dict] = None, music: T.Optional[dict] = None) -> dict:
    """The `json_data.json` sidecar of SonicSet.py:108-136.  Each source: {"audio": [...], "start_end_points":
    [...], "words": [...]} -> keys source1..sourceN, noise, music."""
    out = {"source%d" % (i + 1): dict(audio=s["audio"], start_end_points=s["start_end_points"], words=s.get("words", []))
           for i, s in enumerate(sources)}
    if noise is not None:
        out["noise"] = dict(audio=noise["audio"], start_end_points=noise["start_end_points"])
    if music is not None:
        out["music"] = dict(audio=music["audio"], start_end_points=music["start_end_points"])
    return out


class SceneWriter:
    """Background writer: `put(dir, name, stem, sr)` returns immediately; the WAV is encoded and written on a
    worker thread so that disk I/O overlaps the next scene's rendering."""

    def __init__(self, workers: int = 2):
        self.q = queue.Queue(maxsize=64)
        self.err = None
        self.threads = [threading.Thread(target=self._run, daemon=True) for _ in range(workers)]
        for t in self.threads:
            t.start()

    def _run(self):
        while True:
            job = self.q.get()
            if job is None:
                self.q.task_done()
                return
            try:
                kind, path, payload, sr = job
                os.makedirs(os.path.dirname(path), exist_ok=True)
                if kind == "wav":
                    write_wav_f32(path, payload, sr)
                else:
                    with open(path, "w") as f:
                        json.dump(payload, f)
            except Exception as e:      # noqa: BLE001
                self.err = e
            finally:
                self.q.task_done()

    def put_wav(self, out_dir: str, name: str, stem: np.ndarray, sample_rate: int):
        self.q.put(("wav", os.path.join(out_dir, name), np.array(stem, dtype=np.float32, copy=True), sample_rate))

    def put_json(self, out_dir: str, payload: dict, name: str = "json_data.json"):
        self.q.put(("json", os.path.join(out_dir, name), payload, 0))

    def close(self):
        self.q.join()
        for _ in self.threads:
            self.q.put(None)
        for t in self.threads:
            t.join()
        if self.err is not None:
            raise self.err


def save_scene(writer: SceneWriter, out_dir: str, moving_stems, static_stems, sample_rate: int, sidecar: T.Optional[dict] = None):
    """File names of SonicSet.py:102-106: moving_audio_{i}.wav, noise_audio.wav, music_audio.wav."""
    for i, s in enumerate(moving_stems):
        writer.put_wav(out_dir, "moving_audio_%d.wav" % (i + 1), s, sample_rate)
    for name, s in zip(("noise_audio.wav", "music_audio.wav"), static_stems):
        writer.put_wav(out_dir, name, s, sample_rate)
    if sidecar is not None:
        writer.put_json(out_dir, sidecar)
