timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
run() { timeout 300 python bench.py --steps 20 --warmup 3 --utterances 16 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.readlines()[-1]); r=d['roofline']
print(sys.argv[1], 'value %.0f ms/step %.3f k_render %.1f us/launch k_spectra %.1f e2e %.0f' % (d['value'], d['ms_per_step'], 1e3*r['kernel_ms'], 1e3*r['k_prepare_ms'], d['e2e']['value']))" "$1"; }
run base
SS_EXTRA_NVCC="-DSS_RENDER_MINB=1" python -c "from sonicsim_b200 import _lib; _lib.build(force=True, verbose=True)" 2>&1 | grep -A2 k_render | grep -E "registers|spill"
run minb1
python -c "from sonicsim_b200 import _lib; _lib.build(force=True)"
