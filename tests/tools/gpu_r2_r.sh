set -x
timeout 900 compute-sanitizer --tool memcheck --print-limit 1 python -m pytest tests/test_gpu_parity.py -x -q -m gpu --tb=line -k "418 or 1020" > gpurun_out/sanitizer_fix.log 2>&1
grep "Invalid\|     at \|ERROR SUMMARY\|passed\|failed" gpurun_out/sanitizer_fix.log | head -6
for i in 1 2; do timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu --tb=line 2>&1 | tail -2; done
timeout 900 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r2_r_bench.json 2> gpurun_out/r2_r_bench.err
python - <<PY
import json
d=json.load(open('gpurun_out/r2_r_bench.json'))
print("ms/step", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], "phases", {k: round(v,2) for k,v in d["phases_ms"].items()}, d["parity"]["ok"])
PY
