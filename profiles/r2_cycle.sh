#!/bin/bash
# round-2 measurement cycle: parity tests, bench line, one full ncu capture of the render kernel (7 sources / launch)
# usage: bash profiles/r2_cycle.sh <tag> [kernel regex]
TAG=${1:-r2}; KRE=${2:-k_render}
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py tests/test_gpu_dropin.py -m gpu -x -q 2>&1 | grep -E "passed|failed|error|Error" | tail -4)
timeout 300 python bench.py --steps 20 --warmup 3 --utterances 16 --no-cpu-baseline 2>gpurun_out/bench_$TAG.err | tee gpurun_out/bench_$TAG.json | python -c "
import sys,json
d=json.loads(sys.stdin.readlines()[-1]); r=d['roofline']
print(sys.argv[1], 'value %.0f ms/step %.3f k_render %.1f k_prepare %.1f frac %.4f e2e %.0f' % (d['value'], d['ms_per_step'], 1e3*r['kernel_ms'], 1e3*r['k_prepare_ms'], r['frac'], d['e2e']['value']))" "$TAG"
tail -2 gpurun_out/bench_$TAG.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$KRE -s 2 -c 1 -o gpurun_out/prof_render_$TAG python bench.py --steps 1 --warmup 3 --utterances 7 --inner 1 --no-cpu-baseline > gpurun_out/ncu_render_$TAG.log 2>&1
ls gpurun_out | tr '\n' ' '
