#!/bin/bash
# Round-end validation on one B200 (run through gpurun): GPU tests, sanitizers, every bench workload, the
# reference arm.  Outputs under gpurun_out/<tag>_*; summarised into profiles/ by hand.
tag=${1:-r02z}
out=gpurun_out
if [ "$2" != "sanitizers-only" ]; then python -m pytest tests -q -m gpu 2>&1 | tail -4 > $out/${tag}_pytest_gpu.txt; fi
{
  echo "== compute-sanitizer --tool memcheck, product build"
  timeout 600 compute-sanitizer --tool memcheck python tools/sanitize_k1_k3.py 2>&1 | grep -E "sanitize|ERROR SUMMARY|COMPUTE-SANITIZER|Invalid|illegal"
  timeout 600 compute-sanitizer --tool memcheck python tools/sanitize_k2.py 2>&1 | grep -E "sanitize|ERROR SUMMARY|COMPUTE-SANITIZER|Invalid|illegal"
  echo "== compute-sanitizer --tool memcheck, product build, PYTORCH_NO_CUDA_MEMORY_CACHING=1 (every tensor its own cudaMalloc: exact bounds)"
  PYTORCH_NO_CUDA_MEMORY_CACHING=1 timeout 900 compute-sanitizer --tool memcheck python tools/sanitize_k1_k3.py 2>&1 | grep -E "sanitize|ERROR SUMMARY|COMPUTE-SANITIZER|Invalid|illegal"
  echo "== compute-sanitizer --tool synccheck, product build"
  timeout 600 compute-sanitizer --tool synccheck python tools/sanitize_k1_k3.py 2>&1 | grep -E "sanitize|ERROR SUMMARY|COMPUTE-SANITIZER|illegal"
  timeout 600 compute-sanitizer --tool synccheck python tools/sanitize_k2.py 2>&1 | grep -E "sanitize|ERROR SUMMARY|COMPUTE-SANITIZER|illegal"
  echo "== compute-sanitizer --tool synccheck, product build, PYTORCH_NO_CUDA_MEMORY_CACHING=1"
  PYTORCH_NO_CUDA_MEMORY_CACHING=1 timeout 600 compute-sanitizer --tool synccheck python tools/sanitize_k1_k3.py 2>&1 | grep -E "sanitize|ERROR SUMMARY|COMPUTE-SANITIZER|illegal"
  echo "== compute-sanitizer --tool racecheck, all-lanes-arrive build (python values_b200/build.py --sanitize)"
  VALUES_B200_LIB=$PWD/values_b200/lib_sanitize/libvalues_b200.so timeout 900 compute-sanitizer --tool racecheck python tools/sanitize_k1_k3.py 2>&1 | grep -E "sanitize|ERROR SUMMARY|RACECHECK SUMMARY|COMPUTE-SANITIZER|hazard" | head -20
  echo "== compute-sanitizer --tool racecheck, all-lanes-arrive build, PYTORCH_NO_CUDA_MEMORY_CACHING=1"
  PYTORCH_NO_CUDA_MEMORY_CACHING=1 VALUES_B200_LIB=$PWD/values_b200/lib_sanitize/libvalues_b200.so timeout 900 compute-sanitizer --tool racecheck python tools/sanitize_k1_k3.py 2>&1 | grep -E "sanitize|ERROR SUMMARY|RACECHECK SUMMARY|COMPUTE-SANITIZER|hazard|illegal" | head -20
  VALUES_B200_LIB=$PWD/values_b200/lib_sanitize/libvalues_b200.so timeout 900 compute-sanitizer --tool racecheck python tools/sanitize_k2.py 2>&1 | grep -E "sanitize|ERROR SUMMARY|RACECHECK SUMMARY|COMPUTE-SANITIZER|hazard|illegal" | head -20
} > $out/${tag}_sanitizer.txt 2>&1
if [ "$2" = "sanitizers-only" ]; then cat $out/${tag}_sanitizer.txt; exit 0; fi
python bench.py --impl reference --steps 3 --warmup 1 > $out/${tag}_bench_cfg5_reference_arm.json 2> $out/${tag}_bench.err
python bench.py --steps 20 --warmup 3 > $out/${tag}_bench_cfg5.json 2>> $out/${tag}_bench.err
: > $out/${tag}_bench_others.jsonl
for w in cfg1 cfg2 cfg3 cfg3gauss cfg4 cfg4bf16 cfg5odd cfg4true; do
  python bench.py --workload $w --steps 20 --warmup 3 >> $out/${tag}_bench_others.jsonl 2>> $out/${tag}_bench.err
done
cat $out/${tag}_pytest_gpu.txt; cat $out/${tag}_sanitizer.txt
python - <<PY
import json
for f in ("$out/${tag}_bench_cfg5.json", "$out/${tag}_bench_others.jsonl"):
    for l in open(f):
        l = l.strip()
        if not l.startswith("{"): continue
        d = json.loads(l); r = d["roofline"]; s = d.get("sustained") or {}
        print(d["config"]["workload"][:28], "| %.3g %s | %.3f ms | frac %.3f pipe %.3f | sustained %.3g frac %s | e2e %.3g | parity %s" % (
            d["value"], d["unit"], d["ms_per_step"], r["frac"], r["pipeline_frac"], s.get("value", 0), s.get("frac"),
            (d.get("e2e") or {}).get("value", 0), (d.get("parity") or {}).get("explained_by_map_tolerance")))
PY
