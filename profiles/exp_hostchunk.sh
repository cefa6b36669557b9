run() { timeout 300 python bench.py --steps 20 --warmup 3 --utterances 16 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.readlines()[-1]); r=d['roofline']
print(sys.argv[1], 'value %.0f ms/step %.3f | e2e %.0f ms/step %.3f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step']))" "$1"; }
for mb in 7 14 28 56; do SS_HOST_CHUNK_MB=$mb run host_chunk_$mb; done
