python -m pytest tests/test_gpu_stats.py -q -m gpu -x 2>&1 | tail -3
python tools/k34_bench.py --reps 5 2>&1 | tail -12
