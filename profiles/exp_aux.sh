run() { timeout 300 python bench.py --steps 20 --warmup 3 --utterances 16 --no-cpu-baseline $2 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.readlines()[-1]); r=d['roofline']
print(sys.argv[1], 'value %.0f ms/step %.3f' % (d['value'], d['ms_per_step']))" "$1"; }
run aux2
run aux2_chunk56 "--chunk-mb 56"
SS_EXTRA_NVCC="-DSS_AUX_STREAMS=3" python -c "from sonicsim_b200 import _lib; _lib.build(force=True)" 2>&1 | tail -1
run aux3
run aux3_chunk56 "--chunk-mb 56"
run aux3_chunk42 "--chunk-mb 42"
python -c "from sonicsim_b200 import _lib; _lib.build(force=True)"
