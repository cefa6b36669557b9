run() { timeout 300 python bench.py --steps 20 --warmup 3 --utterances 16 --no-cpu-baseline $2 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.readlines()[-1]); r=d['roofline']
print(sys.argv[1], 'value %.0f ms/step %.3f k_render %.1f us/launch k_prepare %.1f pairs %d e2e %.0f' % (d['value'], d['ms_per_step'], 1e3*r['kernel_ms'], 1e3*r['k_prepare_ms'], r['launch_pairs_timed'], d['e2e']['value']))" "$1"; }
run two_streams_96
SS_SINGLE_STREAM=1 run one_stream_96
run two_streams_56 "--chunk-mb 56"
run two_streams_160 "--chunk-mb 160"
SS_SINGLE_STREAM=1 run one_stream_160 "--chunk-mb 160"
