set -x
python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -8
python -m pytest tests/test_gpu_sharded.py tests/test_gpu_dropin.py tests/test_gpu_bench_workload.py -x -q -m gpu 2>&1 | tail -4
SQG_OTHER_SORT=1 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "short or seeded or golden" 2>&1 | tail -4
P=${1:-100000000}
python bench.py --pairs $P --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r2_d_bench.json 2> gpurun_out/r2_d_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_d_bench.json'))
print("ms/step", d["ms_per_step"], "phases", {k: round(v,2) for k,v in d["phases_ms"].items()}, "timeline", {k: round(v,2) for k,v in d["host_timeline_ms"].items()})
PY
