"""Pin the loudness oracle against the real pyloudnorm 0.1.1 (ss-2.0.yaml:201), if it can be had.

TEST INFRASTRUCTURE.  pyloudnorm is not vendored with the reference, not installed in this image and there is no
network, so the a5 rows of SURVEY 8 are "parity unpinned" (the oracle restates the published algorithm).  This script
is the one-command attempt to change that on any box:

    python oracle/pin_loudness.py            # tries `import pyloudnorm`, then `pip install pyloudnorm==0.1.1`

If the package imports, the reference's own lufs_norm body (SonicSim_audio.py:68-81: Meter(rate, block_size)
.integrated_loudness + pyln.normalize.loudness) is run on the seeded stems of tests/test_loudness.py and the inputs'
seeds + outputs are written to tests/golden/lufs_norm.npz, which tests/test_loudness.py then checks the oracle
(CPU) and the CUDA path (gpu) against.  Every attempt appends its outcome to oracle/pin_loudness.log.
"""
import datetime
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
LOG = os.path.join(ROOT, "oracle", "pin_loudness.log")
OUT = os.path.join(ROOT, "tests", "golden", "lufs_norm.npz")
CASES = [(7 * 480000 + 2, 480000, 2, 16000), (7 * 100001 + 5, 100001, 5, 16000), (7 * 5000 + 2, 5000, 2, 16000),
         (96005, 96000, 5, 48000), (6401, 6400, 1, 16000)]


def log(msg):
    line = "%s  %s" % (datetime.datetime.now(datetime.timezone.utc).strftime("%Y-%m-%dT%H:%M:%SZ"), msg)
    print(line)
    with open(LOG, "a") as f:
        f.write(line + "\n")


def get_pyloudnorm():
    try:
        import pyloudnorm
        return pyloudnorm
    except ImportError:
        pass
    cmd = [sys.executable, "-m", "pip", "install", "--no-input", "--disable-pip-version-check", "pyloudnorm==0.1.1"]
    try:
        res = subprocess.run(cmd, capture_output=True, text=True, timeout=120)
        tail = (res.stdout + res.stderr).strip().splitlines()[-1:] or [""]
        log("pip install pyloudnorm==0.1.1 -> rc %d: %s" % (res.returncode, tail[0][:200]))
    except Exception as e:      # noqa: BLE001
        log("pip install could not run: %r" % (e,))
    try:
        import pyloudnorm
        return pyloudnorm
    except ImportError:
        return None


def reference_lufs_norm(pyln, data, sr, norm):
    """SonicSim_audio.py:68-81, verbatim in behaviour: returns (normalised, gain)."""
    block_size = 0.4 if data.shape[0] > 0.4 * sr else data.shape[0] / sr
    meter = pyln.Meter(rate=sr, block_size=block_size)
    loudness = meter.integrated_loudness(data)
    if np.isinf(loudness):
        loudness = -40
    norm_data = pyln.normalize.loudness(data, loudness, norm)
    return norm_data, loudness


def main():
    pyln = get_pyloudnorm()
    if pyln is None:
        log("pyloudnorm not importable and not installable here: a5 stays 'parity unpinned' (no golden written)")
        return 1
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_loudness import stems
    out = {"version": np.str_(getattr(pyln, "__version__", "unknown")), "n_cases": np.int64(len(CASES))}
    for k, (seed, N, C, sr) in enumerate(CASES):
        x = stems(seed, N, C, sr)
        y, loud = reference_lufs_norm(pyln, x, sr, -17.0)
        out[f"seed{k}"], out[f"N{k}"], out[f"C{k}"], out[f"sr{k}"] = np.int64(seed), np.int64(N), np.int64(C), np.int64(sr)
        out[f"lufs{k}"], out[f"y{k}"] = np.float64(loud), np.asarray(y, dtype=np.float32)[:: max(1, N // 4096)]   # strided sample of y
    np.savez_compressed(OUT, **out)
    log("pyloudnorm %s imported: wrote %s" % (out["version"], OUT))
    return 0


if __name__ == "__main__":
    sys.exit(main())
