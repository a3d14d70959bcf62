set -x
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3
for v in 0 1; do
SQG_SLOW_IN_TILE=$v timeout 900 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r2_o_bench_$v.json 2> gpurun_out/r2_o_bench_$v.err
python - <<PY
import json
d=json.load(open('gpurun_out/r2_o_bench_$v.json'))
print("in_tile=$v ms/step", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], "phases", {k: round(v,2) for k,v in d["phases_ms"].items()}, d["parity"]["ok"], d["stats"])
PY
done
