set -x
python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -5
P=100000000
timeout 900 python bench.py --pairs $P --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_$P.json 2> gpurun_out/bench_$P.err; python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_$P.json')); print({k:d[k] for k in ('value','ms_per_step','phases_ms','gpu_launches','gen_s','stats')}); print(d['e2e']); print(d['config']); print(d['roofline'])
except Exception as e: print("bench failed", e)
PY
tail -5 gpurun_out/bench_$P.err
bash tests/tools/gpu_ncu_list.sh $P 2>&1 | grep -v "^+"
