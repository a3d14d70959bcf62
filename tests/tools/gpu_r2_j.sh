set -x
python -m pytest tests/test_gpu_parity.py tests/test_gpu_sharded.py -x -q -m gpu 2>&1 | tail -3
SQG_SEED_DENSE_R=300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "seeded or golden" 2>&1 | tail -3
SQG_TIMING=1 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r2_j_bench.json 2> gpurun_out/r2_j_bench.err
grep '\[sqg\]' gpurun_out/r2_j_bench.err | tail -12
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_j_bench.json'))
print("ms/step", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], "phases", {k: round(v,2) for k,v in d["phases_ms"].items()}, "timeline", {k: round(v,2) for k,v in d["host_timeline_ms"].items()}, d["stats"]["raw_edges"], d["parity"]["ok"])
PY
bash tests/tools/gpu_r2_i.sh 2>&1 | grep -v "^+" | tail -12
