python -m pytest tests/test_gpu_parity.py tests/test_configs.py -q -m gpu -x -k "c3 or pipeline or cfg" 2>&1 | tail -2
python tools/k2_bench.py --shape 64,64,64 --maps 768 --paths 0 2>&1 | tail -1
for w in cfg2 cfg5; do python bench.py --workload $w --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read()); r=j['roofline']; s=j['sustained']
print(j['config']['workload'][:12], j['dtype'], 'value %.4g ms %.3f frac %.3f pipe %.3f | sustained %.4g frac %.3f pipe %.3f launches %d' % (j['value'], j['ms_per_step'], r['frac'], r['pipeline_frac'], s['value'], s['frac'], s['pipeline_frac'], j['gpu_launches']))"; done
