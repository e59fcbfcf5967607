python tools/k2_bench.py --shape 64,64,64 --maps 768 --paths 0,5 2>&1 | tail -2
python tools/k2_probe.py cfg2 2>&1 | tail -6 | cut -c1-200
ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:k1_tma|box_|stitch|normalize|map_reduce" -c 60 --csv --log-file gpurun_out/r02o_launches_cfg2.csv python bench.py --workload cfg2 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-sustained > /dev/null 2>&1
python bench.py --workload cfg2 --no-cpu-baseline --no-e2e 2>/dev/null | cut -c1-300
