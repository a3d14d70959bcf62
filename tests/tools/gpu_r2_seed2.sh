set -x
python -m pytest tests/test_gpu_parity.py tests/test_gpu_bench_workload.py tests/test_gpu_sharded.py -x -q -m gpu 2>&1 | tail -5
SQG_SEED_DENSE=2 SQG_SEED_DENSE_R=300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3
SQG_SEED_DENSE=0 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3
P=${1:-100000000}
python bench.py --pairs $P --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r2_seed_bench.json 2> gpurun_out/r2_seed_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_seed_bench.json'))
print("ms/step", d["ms_per_step"], "phases", {k: round(v,2) for k,v in d["phases_ms"].items()}, "timeline", {k: round(v,2) for k,v in d["host_timeline_ms"].items()}, d["config"]["segments"], d["config"]["edges"])
PY
