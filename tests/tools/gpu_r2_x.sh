set -x
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sharded.py tests/test_gpu_configs.py -x -q -m gpu --tb=line 2>&1 | tail -2
timeout 900 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r2_x_bench.json 2> gpurun_out/r2_x_bench.err
python - <<PY
import json
d=json.load(open('gpurun_out/r2_x_bench.json'))
print("ms/step", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], "phases", {k: round(v,2) for k,v in d["phases_ms"].items()}, d["parity"])
PY
