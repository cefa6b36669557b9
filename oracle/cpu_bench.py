"""Timing harness for the CPU reference path (the oracle port, which makes the same scipy calls as
SonicSim_moving.py:86-94 and therefore has the reference's cost).  TEST / BENCH INFRASTRUCTURE:
used only by bench.py's `cpu_baseline` leg and `--impl reference` arm.

One worker process per host core, each rendering whole (utterance, source) units single-threaded
(the reference never sets scipy's `workers`), as SURVEY 8(d)(ii) prescribes.
"""
import os
import time

import numpy as np


def cfg2_source(seed, P=40, C=6, L=4096, N=480000):
    """One moving source of BASELINE configs[1] (SURVEY 8d synthetic inputs)."""
    from oracle import sonicsim_oracle as so
    rng = np.random.default_rng(seed)
    x = so.synth_dry(rng, N)
    h = so.synth_rirs(rng, P, C, L)
    pos = so.synth_path(rng, P)
    np.random.seed(seed % (2 ** 31))
    idx, w = so.setup_dynamic_interp(pos, N)
    return x, h, idx, w


_cache = {}


def _worker(args):
    seed, shape, reps = args[:3]
    n_samples = args[3] if len(args) > 3 else None
    for k in ("OMP_NUM_THREADS", "MKL_NUM_THREADS", "OPENBLAS_NUM_THREADS"):
        os.environ[k] = "1"
    from oracle import sonicsim_oracle as so
    key = (seed,) + tuple(shape)
    if key not in _cache:
        _cache.clear()
        _cache[key] = cfg2_source(seed, *shape)
    x, h, idx, w = _cache[key]
    if n_samples is not None and n_samples < x.shape[0]:      # bounded sample: the first n_samples of the unit
        x, idx, w = x[:n_samples], idx[:n_samples], w[:n_samples]
    t0 = time.perf_counter()
    acc = 0.0
    for _ in range(reps):
        y = so.convolve_moving_receiver(x, h, idx, w)
        acc += float(y[0, -1])
    return time.perf_counter() - t0, acc


def pick_workers(shape=(40, 6, 4096, 480000)):
    """Workers = host cores, capped by memory: one unit needs ~ (P*C*N*4 B) * 3 of temporaries."""
    cores = os.cpu_count() or 1
    try:
        cores = len(os.sched_getaffinity(0))
    except Exception:
        pass
    P, C, L, N = shape
    per = 3.0 * P * C * (N + L) * 4 + 1e9
    try:
        import psutil
        avail = psutil.virtual_memory().available
        cores = max(1, min(cores, int(0.6 * avail / per)))
    except Exception:
        pass
    return cores


class CpuPool:
    """Persistent spawn-pool so that imports and input synthesis stay outside the timed region."""

    def __init__(self, workers=None, shape=(40, 6, 4096, 480000)):
        import multiprocessing as mp
        self.shape = shape
        self.workers = workers or pick_workers(shape)
        self.pool = mp.get_context("spawn").Pool(self.workers)
        # warm-up: import scipy, synthesise each worker's unit (untimed), then one timed full-size batch
        self.pool.map(_worker, [(2000 + i, shape, 1, 4 * shape[2]) for i in range(self.workers)], chunksize=1)
        t0 = time.perf_counter()
        self.pool.map(_worker, [(2000 + i, shape, 1) for i in range(self.workers)], chunksize=1)
        self.t_full = time.perf_counter() - t0

    def run_batch(self, reps=1, n_samples=None):
        """Every worker renders (the first n_samples of) its unit `reps` times.  Returns (wall seconds, units)."""
        t0 = time.perf_counter()
        self.pool.map(_worker, [(2000 + i, self.shape, reps, n_samples) for i in range(self.workers)], chunksize=1)
        return time.perf_counter() - t0, self.workers * reps

    def close(self):
        self.pool.close()
        self.pool.join()


def single_thread_time(shape=(40, 6, 4096, 480000), reps=2):
    """The reference exactly as shipped: one process, one thread."""
    _worker((2000, shape, 1))
    t, _ = _worker((2000, shape, reps))
    return t / reps
