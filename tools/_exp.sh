for cfg in 0 3 4; do for pol in 0 512; do
VALUES_K2B_CFG=$cfg VALUES_K2B_POL=$pol ncu --metrics gpu__time_duration.sum --clock-control none -k regex:box_strip --csv python tools/k2_bench.py --shape 128,128,128 --maps 96 --paths 0 --reps 1 2>&1 | grep box_strip | tail -1 | awk -F, -v p=$pol -v c=$cfg '{print "CFG=" c " POL=" p, $NF}'
done; 
echo -n "CFG=$cfg "; VALUES_K2B_CFG=$cfg python tools/k2_bench.py --shape 128,128,128 --maps 96 --paths 0 2>&1 | tail -1 | cut -c1-100
done
