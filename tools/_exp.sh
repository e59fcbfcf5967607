python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_counts.py tests/test_configs.py tests/test_gpu_formats.py tests/test_gpu_patch.py -q -m gpu -x 2>&1 | tail -6
python tools/k1_bench.py --shape cfg3n8 --variants 0,4 2>&1 | tail -2
python tools/k1_bench.py --shape cfg1 --variants 0,4 2>&1 | tail -2
python tools/k1_bench.py --shape cfg5f64 --variants 0 2>&1 | tail -1
python tools/k34_bench.py --reps 5 2>&1 | head -12
echo "== racecheck with the all-lanes-arrive build"
VALUES_B200_LIB=values_b200/lib_sanitize/libvalues_b200.so timeout 900 compute-sanitizer --tool racecheck python tools/sanitize_k1_k3.py 2>&1 | tail -6
VALUES_B200_LIB=values_b200/lib_sanitize/libvalues_b200.so timeout 600 compute-sanitizer --tool racecheck python tools/sanitize_k2.py 2>&1 | tail -4
echo "== memcheck, product build"
timeout 900 compute-sanitizer --tool memcheck python tools/sanitize_k1_k3.py 2>&1 | tail -4
