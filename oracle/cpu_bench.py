"""Timing harness for the reference's CPU path.  TEST / BENCH INFRASTRUCTURE: used only by bench.py's
`cpu_baseline` leg and `--impl reference` arm.

What is timed: `convolve_moving_receiver` (SonicSim_moving.py:63-96) on whole BASELINE configs[1] sources -
the UNMODIFIED reference module from oracle/_ref when oracle/build_ref.py has put it there (kind "reference"),
otherwise the oracle port, which makes the same scipy calls (kind "port").

One worker process per host core, each rendering whole (utterance, source) units single-threaded (the
reference never sets scipy's `workers`), as SURVEY 8(d)(ii) prescribes.  Every worker has its own command
pipe, so a batch can address exactly the first W workers: the number of concurrently active workers is
calibrated (the path allocates ~1.4 GB of temporaries per unit and stops scaling long before 128 processes),
never the length of the signal.
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


def reference_kind():
    from oracle import build_ref
    return "reference" if build_ref.available() else "port"


def _render_fn():
    """The function under test: the unmodified reference if oracle/_ref holds it, else the oracle port."""
    from oracle import build_ref
    ref = build_ref.load()
    if ref is not None:
        return ref.convolve_moving_receiver, "reference"
    from oracle import sonicsim_oracle as so
    return so.convolve_moving_receiver, "port"


def _worker_main(conn, seed, shape):
    for k in ("OMP_NUM_THREADS", "MKL_NUM_THREADS", "OPENBLAS_NUM_THREADS"):
        os.environ[k] = "1"
    try:
        import torch
        torch.set_num_threads(1)
    except Exception:
        pass
    fn, kind = _render_fn()
    x, h, idx, w = cfg2_source(seed, *shape)
    conn.send(("ready", kind))
    while True:
        msg = conn.recv()
        if msg is None:
            break
        reps, n_samples = msg
        xs, ids, ws = x, idx, w
        if n_samples is not None and n_samples < x.shape[0]:
            xs, ids, ws = x[:n_samples], idx[:n_samples], w[:n_samples]
        t0 = time.perf_counter()
        acc = 0.0
        for _ in range(reps):
            y = fn(xs, h, ids, ws)
            acc += float(np.asarray(y)[0, -1])
        conn.send((time.perf_counter() - t0, acc))
    conn.close()


def host_cores():
    cores = os.cpu_count() or 1
    try:
        cores = len(os.sched_getaffinity(0))
    except Exception:
        pass
    return cores


def pick_workers(shape=(40, 6, 4096, 480000)):
    """Workers = host cores, capped by memory: one unit needs ~ (P*C*N*4 B) * 3 of temporaries."""
    cores = host_cores()
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
    """Persistent spawned workers, one pipe each; imports and input synthesis stay outside every timed region."""

    def __init__(self, workers=None, shape=(40, 6, 4096, 480000)):
        import multiprocessing as mp
        ctx = mp.get_context("spawn")
        self.shape = shape
        self.workers = workers or pick_workers(shape)
        self.conns, self.procs = [], []
        for i in range(self.workers):
            a, b = ctx.Pipe()
            p = ctx.Process(target=_worker_main, args=(b, 2000 + i, shape), daemon=True)
            p.start()
            b.close()
            self.conns.append(a)
            self.procs.append(p)
        self.kind = "port"
        for c in self.conns:
            self.kind = c.recv()[1]
        self.run_batch(1, 4 * shape[2])                       # warm-up: scipy plans, allocator
        self.active = self.workers

    def run_batch(self, reps=1, n_samples=None, active=None):
        """The first `active` workers each render (the first n_samples of) their unit `reps` times, concurrently.
        Returns (wall seconds, units rendered)."""
        n = min(active or self.workers, self.workers)
        t0 = time.perf_counter()
        for c in self.conns[:n]:
            c.send((reps, n_samples))
        for c in self.conns[:n]:
            c.recv()
        return time.perf_counter() - t0, n * reps

    def calibrate(self, budget_s=60.0):
        """Full-size batches at W = cores, cores/2, ...: pick the worker count with the highest throughput.
        Returns [(W, seconds per batch)], best first; self.active is set to the best W."""
        tried, w = [], self.workers
        t_used = 0.0
        while w >= 1:
            t, units = self.run_batch(1, None, w)
            tried.append((w, t))
            t_used += t
            if w == 1 or t_used > budget_s:
                break
            # halving the workers can at best keep the batch time: stop when throughput has clearly dropped
            if len(tried) >= 2 and tried[-1][0] / tried[-1][1] < 0.7 * max(a / b for a, b in tried):
                break
            w //= 2
        tried.sort(key=lambda ab: -(ab[0] / ab[1]))
        self.active = tried[0][0]
        return tried

    def close(self):
        for c in self.conns:
            try:
                c.send(None)
            except Exception:
                pass
        for p in self.procs:
            p.join(timeout=5)
            if p.is_alive():
                p.kill()                                       # exact PID of a process this pool started


def single_thread_time(shape=(40, 6, 4096, 480000), reps=2):
    """The reference exactly as shipped: one process, one thread (run in this process)."""
    fn, _ = _render_fn()
    x, h, idx, w = cfg2_source(2000, *shape)
    fn(x, h, idx, w)
    t0 = time.perf_counter()
    for _ in range(reps):
        fn(x, h, idx, w)
    return (time.perf_counter() - t0) / reps
