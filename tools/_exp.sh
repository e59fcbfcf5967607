python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_counts.py tests/test_configs.py -q -m gpu -x -k "c2 or k1 or masks or full_size or cfg or class_mean" 2>&1 | tail -3
python tools/k1_bench.py --shape cfg3n8 --variants 0 2>&1 | tail -1
python tools/k1_bench.py --shape cfg1 --variants 0 2>&1 | tail -1
python tools/k1_bench.py --shape cfg5f64 --variants 0 2>&1 | tail -1
