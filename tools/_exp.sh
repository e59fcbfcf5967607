python -m pytest tests -q -m gpu 2>&1 | tail -4 > gpurun_out/r02q_pytest_gpu.txt; cat gpurun_out/r02q_pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/r02q_bench_cfg5.json 2> gpurun_out/r02q_bench.err; tail -2 gpurun_out/r02q_bench.err
(time python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r02q_bench_cfg5_reference_arm.json 2>> gpurun_out/r02q_bench.err) 2>&1 | tail -3
rm -f gpurun_out/r02q_bench_others.jsonl
for w in cfg1 cfg2 cfg3 cfg4 cfg4bf16; do python bench.py --workload $w >> gpurun_out/r02q_bench_others.jsonl 2>> gpurun_out/r02q_bench.err; done
wc -l gpurun_out/r02q_bench_others.jsonl
