python -m pytest tests/test_gpu_parity.py tests/test_configs.py tests/test_gpu_formats.py tests/test_gpu_patch.py -q -m gpu -x -k "stitch or cfg3 or save_data or normal or install or carrier" 2>&1 | tail -2
for w in cfg3; do python bench.py --workload $w --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read()); r=j['roofline']; s=j['sustained']
print(j['config']['workload'][:12], j['dtype'], 'value %.4g ms %.3f frac %.3f pipe %.3f | sustained %.4g frac %.3f pipe %.3f launches %d' % (j['value'], j['ms_per_step'], r['frac'], r['pipeline_frac'], s['value'], s['frac'], s['pipeline_frac'], j['gpu_launches']))"; done
ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:k1_tma|stitch|normalize" -c 12 --csv --log-file gpurun_out/r02t_launches_cfg3.csv python bench.py --workload cfg3 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-sustained > /dev/null 2>&1
