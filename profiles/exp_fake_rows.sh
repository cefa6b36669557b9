run() { SS_SINGLE_STREAM=1 timeout 300 python bench.py --steps 20 --warmup 3 --utterances 16 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.readlines()[-1]); r=d['roofline']
print(sys.argv[1], 'value %.0f ms/step %.3f k_render %.1f us/launch k_prepare %.1f pairs %d' % (d['value'], d['ms_per_step'], 1e3*r['kernel_ms'], 1e3*r['k_prepare_ms'], r['launch_pairs_timed']))" "$1"; }
run base
SS_EXTRA_NVCC="-DSS_EXP_FAKE_ROWS" python -c "from sonicsim_b200 import _lib; _lib.build(force=True)"
run fake_rows
python -c "from sonicsim_b200 import _lib; _lib.build(force=True)"
