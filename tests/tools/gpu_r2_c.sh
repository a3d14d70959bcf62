set -x
python -m pytest tests -x -q -m gpu 2>&1 | tail -6
P=${1:-100000000}
SQG_TIMING=1 python bench.py --pairs $P --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r2_c_bench.json 2> gpurun_out/r2_c_bench.err
grep '\[sqg\]' gpurun_out/r2_c_bench.err | tail -24
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_c_bench.json'))
print("ms/step", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], "phases", {k: round(v,2) for k,v in d["phases_ms"].items()}, "timeline", {k: round(v,2) for k,v in d["host_timeline_ms"].items()}, d["config"]["segments"], d["config"]["edges"], d["stats"])
PY
