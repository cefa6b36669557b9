#!/bin/bash
# round-2 evidence: launch list, full captures of k_render_fast / k_prepare (7 cfg2 sources per launch) and of the
# kernels around the convolution, their CUDA-event times, per-config times.  usage: bash profiles/r2_profile.sh <tag>
TAG=${1:-r2}
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --utterances 7 --inner 1 --no-cpu-baseline"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${TAG}_launches.csv $B > gpurun_out/ncu_list.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_render -s 2 -c 1 -o gpurun_out/${TAG}_k_render $B > gpurun_out/ncu_render.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_prepare -s 2 -c 1 -o gpurun_out/${TAG}_k_prepare $B > gpurun_out/ncu_prepare.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:k_kweight|k_loud|k_mix|k_dry|k_rir" -c 14 -o gpurun_out/${TAG}_aux python profiles/time_aux.py 1 > gpurun_out/ncu_aux.log 2>&1
timeout 300 python profiles/time_aux.py 10 2>&1 | tee gpurun_out/${TAG}_aux_times.txt
timeout 600 python profiles/time_configs.py 2>&1 | tee gpurun_out/${TAG}_config_times.txt
ls -la gpurun_out | tail -12
