python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "k1 or c2 or c3_golden or stitch_golden" 2>&1 | tail -2
python tools/k1_bench.py --shape cfg5 --batch 8 2>&1 | tail -1
python tools/k1_bench.py --shape cfg4bf16 2>&1 | tail -1
python tools/k1_bench.py --shape cfg2 --batch 256 2>&1 | tail -1
python tools/k1_bench.py --shape cfg3n8 2>&1 | tail -1
python tools/k2_bench.py --shape 128,128,128 --maps 96 --paths 0 2>&1 | tail -1
python tools/k34_bench.py --reps 5 2>&1 | head -2
for w in cfg5; do python bench.py --workload $w --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read()); r=j['roofline']; s=j['sustained']
print(j['config']['workload'][:12], j['dtype'], 'value %.4g ms %.3f frac %.3f pipe %.3f | sustained %.4g frac %.3f pipe %.3f launches %d' % (j['value'], j['ms_per_step'], r['frac'], r['pipeline_frac'], s['value'], s['frac'], s['pipeline_frac'], j['gpu_launches']))"; done
