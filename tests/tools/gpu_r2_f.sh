set -x
python -m pytest tests/test_gpu_parity.py tests/test_gpu_sharded.py tests/test_gpu_bench_workload.py -x -q -m gpu 2>&1 | tail -4
SQG_SLOW_IN_TILE=0 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "seeded or golden or short" 2>&1 | tail -3
for v in 1 0; do
SQG_SLOW_IN_TILE=$v SQG_TIMING=1 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r2_f_bench_$v.json 2> gpurun_out/r2_f_bench_$v.err
grep '\[sqg\]' gpurun_out/r2_f_bench_$v.err | tail -12
python - <<PY
import json
d=json.load(open('gpurun_out/r2_f_bench_$v.json'))
print("SLOW_IN_TILE=$v ms/step", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], "phases", {k: round(v,2) for k,v in d["phases_ms"].items()}, "timeline", {k: round(v,2) for k,v in d["host_timeline_ms"].items()}, d["stats"]["raw_edges"], d["stats"]["edges_generic_path"], d["parity"])
PY
done
