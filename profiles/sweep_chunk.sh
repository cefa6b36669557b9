for mb in 64 96 128 192; do timeout 300 python bench.py --steps 20 --warmup 3 --utterances 16 --no-cpu-baseline --chunk-mb $mb 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.readlines()[-1]); r=d['roofline']
print('chunk_mb', sys.argv[1], 'value %.0f ms/step %.3f k_render %.1f us/launch k_spectra %.1f pairs %d e2e %.0f' % (d['value'], d['ms_per_step'], 1e3*r['kernel_ms'], 1e3*r['k_prepare_ms'], r['launch_pairs_timed'], d['e2e']['value']))" $mb; done
