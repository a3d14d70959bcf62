set -x
SQG_SEED_BLOCK=256 timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu --tb=line 2>&1 | tail -2
run() {
env $1 timeout 900 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r2_t_bench_$2.json 2> gpurun_out/r2_t_bench_$2.err
python - <<PY
import json
d=json.load(open('gpurun_out/r2_t_bench_$2.json'))
print("$1 ms/step", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], "phases", {k: round(v,2) for k,v in d["phases_ms"].items() if k in ("seed","k_seed_islands","classify","depth_edges")}, d["parity"]["ok"])
PY
}
run SQG_SEED_BLOCK=256 c
