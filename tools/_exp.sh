python -m pytest tests -q -m gpu 2>&1 | tail -3 > gpurun_out/r02v_pytest_gpu.txt; cat gpurun_out/r02v_pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
(for tool in memcheck synccheck; do echo "== compute-sanitizer --tool $tool, product build"; compute-sanitizer --tool $tool python tools/sanitize_k1_k3.py 2>&1 | tail -4; compute-sanitizer --tool $tool python tools/sanitize_k2.py 2>&1 | tail -3; done
echo "== compute-sanitizer --tool racecheck, all-lanes-arrive build (python values_b200/build.py --sanitize)"
VALUES_B200_LIB=values_b200/lib_sanitize/libvalues_b200.so compute-sanitizer --tool racecheck python tools/sanitize_k1_k3.py 2>&1 | tail -4
VALUES_B200_LIB=values_b200/lib_sanitize/libvalues_b200.so compute-sanitizer --tool racecheck python tools/sanitize_k2.py 2>&1 | tail -3
echo "== compute-sanitizer --tool racecheck, product build (lane 0 arrives for its warp after __syncwarp: reported as hazards, see DESIGN.md section 6)"
compute-sanitizer --tool racecheck python tools/sanitize_k1_k3.py 2>&1 | grep -E "RACECHECK SUMMARY|Race reported" | sort | uniq -c | head -5) > gpurun_out/r02v_sanitizer.txt 2>&1; cat gpurun_out/r02v_sanitizer.txt
python bench.py > gpurun_out/r02v_bench_cfg5.json 2> gpurun_out/r02v_bench.err; tail -1 gpurun_out/r02v_bench.err
python bench.py --impl reference > gpurun_out/r02v_bench_cfg5_reference_arm.json 2>> gpurun_out/r02v_bench.err
rm -f gpurun_out/r02v_bench_others.jsonl
for w in cfg1 cfg2 cfg3 cfg4 cfg4bf16; do python bench.py --workload $w >> gpurun_out/r02v_bench_others.jsonl 2>> gpurun_out/r02v_bench.err; done
wc -l gpurun_out/r02v_bench_others.jsonl
