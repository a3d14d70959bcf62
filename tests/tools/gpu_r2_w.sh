set -x
run() {
env $1 timeout 900 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r2_w_bench_$2.json 2> gpurun_out/r2_w_bench_$2.err
python - <<PY
import json
d=json.load(open('gpurun_out/r2_w_bench_$2.json'))
print("$1 ms/step", d["ms_per_step"], "phases", {k: round(v,2) for k,v in d["phases_ms"].items()}, d["parity"]["ok"])
PY
}
run SQG_COV_FORK=seed a
run SQG_COV_FORK=depth b
SQG_COV_FORK=depth timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sharded.py -x -q -m gpu --tb=line 2>&1 | tail -2
