ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:box_" --csv --log-file gpurun_out/r02r_launch_128.csv python tools/k2_case.py 128,128,128 96 >/dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:box_" --csv --log-file gpurun_out/r02r_launch_64.csv python tools/k2_case.py 64,64,64 768 >/dev/null 2>&1
