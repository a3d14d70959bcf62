set -x
python -m pytest tests/test_gpu_configs.py -x -q -m gpu 2>&1 | tail -6
python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3
SQG_TIMING=1 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r2_k_bench.json 2> gpurun_out/r2_k_bench.err
grep '\[sqg\]' gpurun_out/r2_k_bench.err | tail -12
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_k_bench.json'))
print("ms/step", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], "phases", {k: round(v,2) for k,v in d["phases_ms"].items()}, "timeline", {k: round(v,2) for k,v in d["host_timeline_ms"].items()}, d["parity"]["ok"])
PY
