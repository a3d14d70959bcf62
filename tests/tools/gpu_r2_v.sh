set -x
run() {
env $1 timeout 900 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r2_v_bench_$2.json 2> gpurun_out/r2_v_bench_$2.err
python - <<PY
import json
d=json.load(open('gpurun_out/r2_v_bench_$2.json'))
print("$1 ms/step", d["ms_per_step"], "phases", {k: round(v,2) for k,v in d["phases_ms"].items() if k in ("seed","k_seed_islands")}, d["parity"]["ok"], d["stats"]["heavy_islands"])
PY
}
run SQG_HEAVY_SPAN=8192 a
run SQG_HEAVY_SPAN=32768 b
run SQG_HEAVY_SPAN=65536 c
run SQG_HEAVY_SPAN=131072 d
