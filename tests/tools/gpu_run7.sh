set -x
python -m pytest tests -x -q -m gpu 2>&1 | tail -8
python -m pytest tests -x -q -m "not gpu" 2>&1 | tail -3
timeout 1200 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_default.json')); print({k:d[k] for k in ('value','ms_per_step','phases_ms','gpu_launches','cpu_baseline','clocks')}); print(d['e2e']); print(d['roofline'])
except Exception as e: print("bench failed", e)
PY
tail -5 gpurun_out/bench_default.err
timeout 900 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; cat gpurun_out/bench_reference.json | head -c 1500; tail -3 gpurun_out/bench_reference.err
nproc; lscpu | grep "Model name"
