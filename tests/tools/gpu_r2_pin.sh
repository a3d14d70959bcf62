set -x
nproc; free -g | head -2; df -h /dev/shm /tmp | tail -2
python -m pytest tests/test_gpu_parity.py tests/test_gpu_sharded.py tests/test_gpu_dropin.py tests/test_gpu_bench_workload.py -x -q -m gpu 2>&1 | tail -5
SQG_DEVICE_PREPASS=0 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "seeded or golden or short" 2>&1 | tail -3
SQG_TIMING=1 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r2_e_bench.json 2> gpurun_out/r2_e_bench.err
grep '\[sqg\]' gpurun_out/r2_e_bench.err | tail -13
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_e_bench.json'))
print("ms/step", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], "K", d["config"]["blocks_per_record"], "phases", {k: round(v,2) for k,v in d["phases_ms"].items()}, "timeline", {k: round(v,2) for k,v in d["host_timeline_ms"].items()}, d["config"]["segments"], d["config"]["edges"], d["stats"], d["parity"])
PY
python tests/tools/pin_bench_crc.py --out gpurun_out/bench_crc_new.json 2>&1 | tail -12
