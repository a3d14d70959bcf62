set -x
run() {
env $1 timeout 900 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r2_p_bench_$2.json 2> gpurun_out/r2_p_bench_$2.err
python - <<PY
import json
d=json.load(open('gpurun_out/r2_p_bench_$2.json'))
print("$1 ms/step", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], "phases", {k: round(v,2) for k,v in d["phases_ms"].items() if k.startswith("k_") or k=="depth_edges"}, d["parity"]["ok"])
PY
}
run SQG_GENERIC_OCC=7 a
run SQG_GENERIC_OCC=8 b
run SQG_GENERIC_OCC=6 c
