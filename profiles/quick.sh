#!/bin/bash
# quick check of a kernel change on the GPU box: parity tests + one bench line.  usage: bash profiles/quick.sh <tag>
TAG=${1:-q}
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -m gpu -x -q 2>&1 | grep -E "passed|failed|error" | tail -2)
timeout 300 python bench.py --steps 20 --warmup 3 --utterances 16 --no-cpu-baseline 2>/dev/null | tee gpurun_out/bench_$TAG.json | python -c "
import sys,json
d=json.loads(sys.stdin.readlines()[-1]); r=d['roofline']
print(sys.argv[1], 'value %.0f ms/step %.3f k_render %.1f k_prepare %.1f e2e %.0f' % (d['value'], d['ms_per_step'], 1e3*r['kernel_ms'], 1e3*r['k_prepare_ms'], d['e2e']['value']))" "$TAG"
