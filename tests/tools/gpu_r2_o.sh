set -x
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_wire.py -x -q -m gpu 2>&1 | tail -3
run() {
env $1 timeout 900 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r2_o_bench_$2.json 2> gpurun_out/r2_o_bench_$2.err
python - <<PY
import json
d=json.load(open('gpurun_out/r2_o_bench_$2.json'))
print("$1 ms/step", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], "phases", {k: round(v,2) for k,v in d["phases_ms"].items()}, d["parity"]["ok"], d["clocks"])
PY
}
run SQG_DEPTH_OVERLAP=1 a
run SQG_DEPTH_OVERLAP=0 b
run SQG_SLOW_IN_TILE=1 c
