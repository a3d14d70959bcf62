set -x
timeout 900 python -m pytest tests/test_gpu_wire.py -x -q -m gpu 2>&1 | tail -8
timeout 1500 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_m_bench.json 2> gpurun_out/r2_m_bench.err
tail -5 gpurun_out/r2_m_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_m_bench.json'))
print("ms/step", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], "soa", d["e2e_soa"]["ms_per_step"], "pack", d["e2e"]["pack_wire_s_outside_timed_region"], d["e2e"]["input"])
print("phases", {k: round(v,2) for k,v in d["phases_ms"].items()})
print("phases e2e", {k: round(v,2) for k,v in d["phases_ms_e2e"].items()})
print("timeline", {k: round(v,2) for k,v in d["host_timeline_ms"].items()}, d["parity"])
print("roof", d["roofline"]["kernel"], d["roofline"]["frac"], {k:(round(v["ms"],2), v["frac"] and round(v["frac"],3)) for k,v in d["roofline"]["kernels"].items()})
print("cpu", json.dumps(d["cpu_baseline"]))
PY
