# pass-C twiddle table (30 KB per transform) vs the ~28 KB of L1 left beside 2 x 103 KB of shared memory:
# rows >= SS_TWC_STREAM_FROM bypass L1 so that the rest stays resident
run() { timeout 300 python bench.py --steps 20 --warmup 3 --utterances 16 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.readlines()[-1]); r=d['roofline']
print(sys.argv[1], 'value %.0f ms/step %.3f k_render %.1f k_prepare %.1f' % (d['value'], d['ms_per_step'], 1e3*r['kernel_ms'], 1e3*r['k_prepare_ms']))" "$1"; }
run base
for f in 12 8 4 1; do
  SS_EXTRA_NVCC="-DSS_TWC_STREAM_FROM=$f" python -c "from sonicsim_b200 import _lib; _lib.build(force=True)" 2>&1 | tail -1
  run "stream_from=$f"
done
python -c "from sonicsim_b200 import _lib; _lib.build(force=True)"
