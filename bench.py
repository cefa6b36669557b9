#!/usr/bin/env python
"""bench.py - rendered audio seconds per second on BASELINE configs[1]
(2 moving speakers x 6 mics x 40-point trajectory x 30 s @ 16 kHz, L = 4096 pre-baked RIRs).

  python bench.py --gpus N --steps K --warmup W            # this repo (CUDA, one process per GPU)
  python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path on host cores

One step = one pass of the hot path (SonicSim_moving.convolve_moving_receiver for both speakers)
over a batch of U utterances per GPU.  `value` = whole-job mixture-seconds per second with inputs
resident in HBM; `e2e` = the same through ss_render_host with pinned HOST buffers (H2D + D2H inside
the timed region).  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "rendered_audio_seconds_per_second"
UNIT = "audio-s/s"
CFG = dict(P=40, C=6, L=4096, N=480000, sr=16000, speakers=2)


def alg_bytes_moving(N, P, C, L):
    """algorithmic HBM bytes of one moving source = 4 (N + P C L + C N)."""
    return 4.0 * (N + P * C * L + C * N)


def workload_name(U, inner=1):
    return ("cfg2: 2 moving speakers x 6-mic x 40-point trajectory x 30 s @16 kHz, L=4096 taps; "
            "%d utterances per GPU per pass x %d passes = %d utterances per GPU per step" % (U, inner, U * inner))


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples SM clock / throttle reasons of one GPU through NVML every ~10 ms."""

    def __init__(self, cuda_index):
        self.samples = []
        self.ok = False
        self._stop = threading.Event()
        try:
            import pynvml
            import torch
            self.nv = pynvml
            pynvml.nvmlInit()
            h = None
            try:
                uuid = str(torch.cuda.get_device_properties(cuda_index).uuid)
                h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
            except Exception:
                h = pynvml.nvmlDeviceGetHandleByIndex(cuda_index)
            self.h = h
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:          # noqa: BLE001
            self.err = repr(e)

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                pw = nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0
                self.samples.append((time.perf_counter(), sm, rs, pw))
            except Exception:
                pass
            time.sleep(0.01)

    def start(self):
        if self.ok:
            self.th = threading.Thread(target=self._run, daemon=True)
            self.th.start()

    def stop(self):
        self._stop.set()
        if self.ok:
            self.th.join(timeout=2)

    def summary(self, windows):
        if not self.ok:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable: " + getattr(self, "err", "")]}
        nv = self.nv
        sel = [s for s in self.samples if any(a <= s[0] <= b for a, b in windows)] or self.samples
        names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4),
                 "hw_power_brake": getattr(nv, "nvmlClocksThrottleReasonHwPowerBrakeSlowdown", 0x80)}
        reasons = sorted(k for k, bit in names.items() if any(s[2] & bit for s in sel))
        return {"sm_mhz": float(np.median([s[1] for s in sel])) if sel else None,
                "sm_max_mhz": float(self.max_sm), "reasons": reasons, "samples": len(sel),
                "power_w_max": max([s[3] for s in sel]) if sel else None}


# ----------------------------------------------------------------------------- reference arm
def run_reference(args, rank, world):
    """The reference's own CPU implementation of the path on the host cores: the unmodified SonicSim_moving module
    from oracle/_ref (kind "reference") when build() has copied it there, else the oracle port (kind "port").
    One step = one batch of W whole cfg2 sources (30 s each, never truncated), one per worker process; W is the
    worker count with the best measured throughput whose K + W_up batches fit the time budget."""
    if rank != 0:
        return
    from oracle import cpu_bench
    shape = (CFG["P"], CFG["C"], CFG["L"], CFG["N"])
    pool = cpu_bench.CpuPool(shape=shape)
    cal = pool.calibrate(budget_s=60.0)               # [(workers, seconds per full batch)], best throughput first
    n_batches = args.steps + max(0, args.warmup - 1)
    budget_s = 170.0
    fit = [c for c in cal if c[1] * n_batches <= budget_s]
    w_use, t_batch = fit[0] if fit else min(cal, key=lambda c: c[1])
    for _ in range(max(0, args.warmup - 1)):
        pool.run_batch(1, None, w_use)
    t_tot, units, per_step = 0.0, 0, []
    for _ in range(args.steps):
        t, u = pool.run_batch(1, None, w_use)
        t_tot += t
        units += u
        per_step.append(t)
    kind = pool.kind
    pool.close()
    unit_s = CFG["N"] / CFG["sr"] / CFG["speakers"]      # mixture-seconds one source stands for
    value = units * unit_s / t_tot
    sample = ("%d steps x %d whole cfg2 sources (%.0f s each, untruncated; one per worker process, 1 thread each) = %.0f "
              "mixture-seconds; worker count calibrated on full batches: %s" % (args.steps, w_use, CFG["N"] / CFG["sr"], units * unit_s,
              ", ".join("%d workers %.1f s" % c for c in sorted(cal, key=lambda c: -c[0]))))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_tot / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": {"workload": workload_name(args.utterances, args.inner)},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": w_use, "kind": kind, "sample": sample,
                             "host_cpu_count": os.cpu_count(), "threads_per_worker": 1,
                             "median_step_value": w_use * unit_s / float(np.median(per_step)),
                             "note": ("kind=reference: oracle/_ref/SonicSim_moving.py is the unmodified reference module "
                                      "(oracle/build_ref.py copies it); kind=port: the oracle's restatement with the same scipy "
                                      "calls (SonicSim_moving.py:86-94).  cores = worker processes active per step, chosen for "
                                      "the best throughput that fits the time budget")},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- our arm
def make_inputs(U, rank):
    """U utterances x 2 speakers of cfg2; seeds 1000*cfg + item."""
    from oracle import sonicsim_oracle as so          # synthetic-input generators only
    from sonicsim_b200 import render
    items = []
    for u in range(U * CFG["speakers"]):
        seed = 2000 + rank * 100000 + u
        rng = np.random.default_rng(seed)
        x = so.synth_dry(rng, CFG["N"])
        h = so.synth_rirs(rng, CFG["P"], CFG["C"], CFG["L"])
        pos = so.synth_path(rng, CFG["P"])
        np.random.seed(seed % (2 ** 31))
        b = render.trajectory_bounds(pos, CFG["N"])
        items.append((x, h, b))
    return items


def run_ours(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    from sonicsim_b200 import render, shard
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    numa_node = shard.bind_to_gpu_numa(local_rank) if world > 1 else None      # keep pinned buffers socket-local
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    U, K, W, I = args.utterances, args.steps, args.warmup, max(1, args.inner)
    R = render.Renderer(local_rank)
    if args.chunk_mb:
        R.set_chunk_bytes(args.chunk_mb << 20)
    items = make_inputs(U, rank)
    n_src = len(items)
    C, N = CFG["C"], CFG["N"]

    # ---- device-resident arm
    d_srcs = [render.MovingSource(torch.from_numpy(x).to(dev), torch.from_numpy(h).to(dev), torch.from_numpy(b).to(dev), b)
              for x, h, b in items]
    d_outs = [torch.empty((C, N), dtype=torch.float32, device=dev) for _ in range(n_src)]
    # one step = I passes over the batch of U utterances (a step of a single pass is 0.7 ms: the timed region of the
    # driver's 20 steps would be 15 ms, in which one host hiccup on one of 8 ranks decides the scaling efficiency)
    pass_alg = sum(alg_bytes_moving(N, CFG["P"], C, CFG["L"]) for _ in items)
    step_alg = pass_alg * I
    step_audio = U * I * N / CFG["sr"]

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    sampler.start()
    dplan = R.plan_device(d_srcs, d_outs)           # batch bound to its device tensors once; run() = one C-ABI call
    for _ in range(W):
        for _ in range(I):
            dplan.run()
    barrier()
    R.reset_stats()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_w0 = time.perf_counter()
    e0.record()
    for _ in range(K):
        for _ in range(I):
            dplan.run()                               # one C-ABI call = one CUDA graph launch (10 kernels)
    e1.record()
    t_issue = time.perf_counter() - t_w0          # host time to enqueue the K steps (launch-bound if ~ device time)
    barrier()
    t_w1 = time.perf_counter()
    launches = R.launch_count()
    dev_s = e0.elapsed_time(e1) / 1e3

    # ---- per-kernel timing (the same pass again, CUDA events around every launch)
    n_prof = min(K * I, 24)
    R.set_profiling(True)
    for _ in range(n_prof):
        dplan.run()
    torch.cuda.synchronize()
    ms_spec, ms_rend, n_pairs = R.get_profile()
    R.set_profiling(False)

    # ---- end-to-end arm: pinned host buffers through ss_render_host
    h_items = [(torch.from_numpy(x).pin_memory(), torch.from_numpy(h).pin_memory(), b) for x, h, b in items]
    h_srcs = [render.MovingSource(x.numpy(), h.numpy(), b) for x, h, b in h_items]
    h_outs_t = [torch.empty((C, N), dtype=torch.float32).pin_memory() for _ in range(n_src)]
    h_outs = [t.numpy() for t in h_outs_t]
    plan = R.plan_host(h_srcs, h_outs)            # batch validated and bound to its pinned buffers once
    for _ in range(max(1, min(W, 2))):
        plan.run()
    barrier()
    t_e0 = time.perf_counter()
    for _ in range(K):
        for _ in range(I):
            plan.run()                            # one ss_render_host call: H2D -> render -> D2H
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t_e0
    barrier()
    t_e1 = time.perf_counter()
    # ---- secondary: the same end-to-end pass with every stem loudness-normalised on the device before its copy out
    # (SonicSet.py:97-101 calls get_lufs_norm_audio on every rendered stem)
    plan_l = R.plan_host(h_srcs, h_outs, lufs_targets=[-17.0] * n_src, sr=CFG["sr"])
    plan_l.run()
    barrier()
    n_l = max(2, min(K * I, 8))
    t_l0 = time.perf_counter()
    for _ in range(n_l):
        plan_l.run()
    torch.cuda.synchronize()
    lufs_s = (time.perf_counter() - t_l0) / n_l
    lufs_meas = plan_l.loudness()[0]
    barrier()
    # ---- copy-only floor of the e2e arm: the same bytes in both directions over PCIe, nothing else, on every rank at
    # the same time (at N >= 4 the GPUs share PCIe switches / host memory and the floor itself rises)
    n_floor = max(2, min(K * I, 12))
    s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    def copy_pass():
        with torch.cuda.stream(s_in):
            for (hx, hh, _), ds in zip(h_items, d_srcs):
                ds.dry.copy_(hx, non_blocking=True)
                ds.rirs.copy_(hh, non_blocking=True)
        with torch.cuda.stream(s_out):
            for ho, do in zip(h_outs_t, d_outs):
                ho.copy_(do, non_blocking=True)
    copy_pass()
    barrier()
    t_f0 = time.perf_counter()
    for _ in range(n_floor):
        copy_pass()
    torch.cuda.synchronize()
    floor_s = (time.perf_counter() - t_f0) / n_floor
    barrier()
    sampler.stop()
    # the device arm and the host arm must agree bit for bit (same kernels)
    same = bool(np.array_equal(h_outs[0], d_outs[0].cpu().numpy()))
    checksum = float(np.abs(h_outs[-1]).sum())          # the D2H'd result is really read

    # ---- max over ranks (device time), counters all-gather (the path's only collective)
    tt = torch.tensor([dev_s, e2e_s, floor_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    dev_max, e2e_max, floor_max = float(tt[0]), float(tt[1]), float(tt[2])
    counters = shard.gather_counters(step_audio * K, dev_s, step_alg * K, device=dev)
    issue_all = shard.gather_counters(t_issue, e2e_s, 0.0, device=dev)
    total_audio = float(counters[:, 0].sum())
    value = total_audio / dev_max
    e2e_value = total_audio / e2e_max

    if rank == 0:
        peaks, peak_src = {}, "fallback (B200_PROFILING.md: 6650 GB/s)"
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
            peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        k_render_s = ms_rend / 1e3 / max(n_pairs, 1)
        k_spec_s = ms_spec / 1e3 / max(n_pairs, 1)
        alg_per_launch = pass_alg * n_prof / max(n_pairs, 1)
        achieved = alg_per_launch / k_render_s / 1e9 if k_render_s > 0 else 0.0
        achieved_path = step_alg * K / dev_s / 1e9              # whole hot path over the timed region itself
        traffic, traffic_note, ncu_extra, traffic_parts = None, "no ncu capture found (profiles/ncu_traffic.json)", None, None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
            ncu_extra = {k: tj["k_render"].get(k) for k in ("issue_slots_busy_pct", "fma_pipe_active_pct", "dram_pct_of_peak")}
            mb = lambda k: (tj[k]["dram_read_mb"] + tj[k]["dram_write_mb"]) * 1e6 / tj["sources_per_launch"]
            scale = n_src * n_prof / max(n_pairs, 1)
            traffic = (mb("k_render") + mb("k_prepare")) * scale
            traffic_parts = {"k_render": mb("k_render") * scale, "k_prepare": mb("k_prepare") * scale}
            traffic_note = ("dram__bytes_read.sum + dram__bytes_write.sum of one k_prepare + k_render launch pair from `ncu --set "
                            "full` (%s; cold L2 per replay: the spectra k_prepare writes and k_render reads make an HBM round "
                            "trip there that they do not make in the timed run), scaled to this launch size; compare with "
                            "alg_bytes_per_launch" % tj["tag"])
        except Exception:
            pass
        # instruction-issue view of the same launch (the kernel's actual limiter): warp instructions of one launch
        # from the ncu capture, scaled to this launch size, over the issue slots of its measured duration
        clocks = sampler.summary([(t_w0, t_w1), (t_e0, t_e1)])
        issue = None
        try:
            sm_count = torch.cuda.get_device_properties(dev).multi_processor_count
            mhz = float(clocks.get("sm_mhz") or clocks.get("sm_max_mhz") or 0.0)
            winst = tj["k_render"]["warp_instructions"] / tj["sources_per_launch"] * (n_src * n_prof / max(n_pairs, 1))
            slots = k_render_s * sm_count * 4 * mhz * 1e6            # 4 schedulers per SM, one warp instruction per cycle each
            if slots > 0:
                issue = {"warp_instructions_per_launch": winst, "issue_slots": slots, "frac": winst / slots,
                         "note": "warp instructions (ncu, %s) / (kernel_ms x %d SMs x 4 schedulers x %.0f MHz)"
                                 % (tj["tag"], sm_count, mhz)}
        except Exception:
            pass
        # L1 / shared-memory data-pipe view (the other resource the kernel runs against): bytes one
        # transform moves through an SM's 128 B/clk data pipe - staged spectra written by the copy engine (Hp, Hq every
        # transform, X once per block = every C-th) and read once (96 KB), two exchanges through the FFT buffer
        # (2 x (64 + 64) KB), inter-pass twiddles (2 x 4 table rows x 8 B x 256 threads), the (4096,) float32 output row
        pipe = None
        try:
            n_tr = sum(int(np.ceil(np.diff(b) / 4096.0).sum()) * C for _, _, b in items) * n_prof / max(n_pairs, 1)
            per_tr = (2.0 + 1.0 / C) * 32 * 1024 + 96 * 1024 + 2 * 128 * 1024 + 2 * 4 * 8 * 256 + 4096 * 4
            pk = sm_count * 128 * mhz * 1e6
            pipe = {"transforms_per_launch": n_tr, "bytes_per_transform": per_tr, "achieved": n_tr * per_tr / k_render_s / 1e9,
                    "peak": pk / 1e9, "unit": "GB/s", "frac": n_tr * per_tr / k_render_s / pk,
                    "note": "algorithmic bytes through the SMs' L1 / shared-memory data pipes (%d SMs x 128 B/clk x %.0f MHz); "
                            "bank conflicts and the copy engine's own arbitration come on top" % (sm_count, mhz)}
        except Exception:
            pass
        in_b = sum(x.nbytes + h.nbytes + b.nbytes for x, h, b in items) * I
        out_b = n_src * C * N * 4 * I
        scene_lufs = {"value": U * N / CFG["sr"] / lufs_s, "unit": UNIT, "ms_per_pass": 1e3 * lufs_s,
                      "extra_ms_per_pass_vs_e2e": 1e3 * (lufs_s - e2e_s / K / I), "rank": 0,
                      "first_stem_lufs_before": lufs_meas[0], "first_stem_gain": lufs_meas[1],
                      "what": "e2e pass (rank 0) with lufs_norm fused behind the render: K-weighting (3 exact passes), gating, "
                              "gain and in-place scale of every (6, 480000) stem on the device before its D2H copy; "
                              "Renderer.plan_host(..., lufs_targets).run() -> ss_render_host_ex"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": 1e3 * dev_max / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(U, I), "utterances_per_gpu_per_step": U * I,
                       "sources_per_gpu_per_step": n_src * I, "passes_per_step": I,
                       "l2": "inputs %.0f MB + outputs %.0f MB per pass are larger than the 126 MB L2 (no flush needed)"
                             % (in_b / I / 1e6, out_b / I / 1e6),
                       "device_arm": "Renderer.plan_device(...).run() -> ss_plan_run: %s" %
                                     ("one CUDA graph launch per pass" if dplan.is_graph() else "direct launches (no graph)"),
                       "parallelism": "units sharded across %d rank(s), no data-path collective" % world,
                       "numa_node_rank0": numa_node,
                       "per_rank_ms_per_step": [round(1e3 * float(t) / K, 4) for t in counters[:, 1]],
                       "per_rank_host_issue_ms_per_step": [round(1e3 * float(t) / K, 4) for t in issue_all[:, 0]],
                       "per_rank_e2e_ms_per_step": [round(1e3 * float(t) / K, 3) for t in issue_all[:, 1]]},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(in_b), "d2h_bytes_per_step": int(out_b),
                    "ms_per_step": 1e3 * e2e_max / K, "ms_per_pass": 1e3 * e2e_max / K / I,
                    "copy_floor_ms_per_pass": 1e3 * floor_max,
                    "frac_of_copy_floor": floor_max / (e2e_max / K / I),
                    "copy_floor_note": "pinned H2D of a pass's inputs and D2H of its outputs on two streams, no kernels, all "
                                       "ranks at once (max over ranks); the e2e arm cannot be faster than this on this box",
                    "api": "sonicsim_b200.render.Renderer.plan_host(...).run() -> ss_render_host",
                    "bit_identical_to_device_arm": same, "checksum": checksum},
            "scene_lufs": scene_lufs,
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": "k_render_fast", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "traffic_parts": traffic_parts, "traffic_note": traffic_note, "peak_source": peak_src,
                         "limiter": "instruction issue (~0.6 of the slots) and the SMs' L1 / shared-memory data pipe (~0.65) together, not DRAM (see DESIGN.md section 4)",
                         "ncu_k_render": ncu_extra, "issue": issue, "l1_data_pipe": pipe,
                         "alg_bytes_per_launch": alg_per_launch, "kernel_ms": 1e3 * k_render_s,
                         "k_prepare_ms": 1e3 * k_spec_s, "path_achieved": achieved_path,
                         "path_frac": achieved_path / peak, "launch_pairs_timed": int(n_pairs),
                         "timing": "kernel_ms / k_prepare_ms: CUDA events recorded by the library on the launching stream "
                                   "around every launch, K steps repeated right after the timed region with the chunk overlap "
                                   "(three internal streams) switched off so that each kernel runs alone; per launch position "
                                   "the median over the K steps is used; path_achieved: "
                                   "algorithmic bytes of the timed region / its device time (overlap on)"},
        }
        if world == 1 and not args.no_cpu_baseline:
            from oracle import cpu_bench
            shape = (CFG["P"], CFG["C"], CFG["L"], CFG["N"])
            t1 = cpu_bench.single_thread_time(shape, reps=1)
            pool = cpu_bench.CpuPool(shape=shape)
            cal = pool.calibrate(budget_s=48.0)          # bounded sample: full batches at cores, cores / 2, ...
            w_use, t = cal[0]
            kind = pool.kind
            pool.close()
            unit_s = N / CFG["sr"] / CFG["speakers"]
            line["cpu_baseline"] = {
                "value": w_use * unit_s / t, "unit": UNIT, "cores": w_use, "kind": kind,
                "sample": "%d whole cfg2 sources (one per worker process, 1 thread each) = %.0f mixture-seconds in %.1f s; "
                          "best of the worker counts tried: %s" % (w_use, w_use * unit_s, t,
                          ", ".join("%d workers %.1f s" % c for c in sorted(cal, key=lambda c: -c[0]))),
                "single_thread_value": unit_s / t1,
                "host_cpu_count": os.cpu_count()}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--utterances", type=int, default=16, help="utterances per GPU per pass")
    ap.add_argument("--inner", type=int, default=32, help="passes over the batch per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--chunk-mb", type=int, default=0, help="override the library's L2-sized spectra budget")
    ap.add_argument("--tiny", action="store_true", help="(tests only) shrink the workload to seconds of CPU time")
    args = ap.parse_args()
    if args.tiny:
        CFG.update(P=6, C=2, L=512, N=24000)
        args.inner = 1
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        if args.warmup < 3:
            args.warmup = 3
        run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
