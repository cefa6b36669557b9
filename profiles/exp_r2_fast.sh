# round 2: knobs of k_render_fast (split barriers, hoisted twiddle loads, last-warp copy issue ...)
run() { timeout 300 python bench.py --steps 20 --warmup 3 --utterances 16 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.readlines()[-1]); r=d['roofline']
print(sys.argv[1], 'value %.0f ms/step %.3f k_render %.1f k_prepare %.1f' % (d['value'], d['ms_per_step'], 1e3*r['kernel_ms'], 1e3*r['k_prepare_ms']))" "$1"; }
build() { SS_EXTRA_NVCC="$1" python -c "from sonicsim_b200 import _lib; _lib.build(force=True)" 2>&1 | tail -1; }
run base
for f in "$@"; do
  build "$f"; run "[$f]"
done
build ""
