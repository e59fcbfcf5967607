set -x
K='regex:k1_|box_|stitch_|normalize|map_reduce|radix|patch_'
ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 200 --csv --log-file gpurun_out/r02u_launches_cfg5.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-sustained > gpurun_out/r02u_b.log 2>&1
for w in cfg1 cfg2 cfg3 cfg4; do ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 200 --csv --log-file gpurun_out/r02u_launches_$w.csv python bench.py --workload $w --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-sustained --no-sweep >> gpurun_out/r02u_b.log 2>&1; done
ncu --set full --clock-control none --import-source on -k regex:k1_tma -c 1 -o gpurun_out/r02u_k1_f64 python tools/k1_bench.py --shape cfg3n8 --reps 1 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:box_strip -c 1 -o gpurun_out/r02u_k2b_strip python tools/k2_case.py 128,128,128 96 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:box_strip -c 1 -o gpurun_out/r02u_k2b_strip_2d python tools/k2_case.py 1024,2048 18 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:box_march -c 1 -o gpurun_out/r02u_k2b_pass1 python tools/k2_case.py 128,128,128 96 > /dev/null 2>&1
ls gpurun_out/r02u*
