#!/bin/bash
# One measurement cycle on the GPU box: parity tests, bench (device + e2e), ncu launch list and a
# full capture of the dominant kernels.  usage: bash profiles/gpu_cycle.sh <tag> [utterances]
TAG=${1:-vX}; U=${2:-16}
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4)
timeout 600 python bench.py --steps 20 --warmup 3 --utterances $U --no-cpu-baseline > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_$TAG.json").read().strip().splitlines()[-1])
r=d["roofline"]
print("BENCH $TAG value %.0f ms/step %.3f | k_render %.1f us k_prepare %.1f us frac %.4f path_frac %.4f | e2e %.0f (%.2f ms) | clocks %s" % (
  d["value"], d["ms_per_step"], 1e3*r["kernel_ms"], 1e3*r["k_prepare_ms"], r["frac"], r["path_frac"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["clocks"]))
PY
tail -3 gpurun_out/bench_$TAG.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --utterances 7 --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_render -s 2 -c 1 -o gpurun_out/prof_render_$TAG python bench.py --steps 1 --warmup 3 --utterances 7 --no-cpu-baseline > gpurun_out/ncu_render.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_prepare -s 2 -c 1 -o gpurun_out/prof_prepare_$TAG python bench.py --steps 1 --warmup 3 --utterances 7 --no-cpu-baseline > gpurun_out/ncu_spectra.log 2>&1
ls gpurun_out | tr '\n' ' '
