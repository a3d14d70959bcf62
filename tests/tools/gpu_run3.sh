set -x
python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -5
for P in 10000000; do
timeout 600 python bench.py --pairs $P --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/bench_$P.json 2> gpurun_out/bench_$P.err; python - <<PY
import json
d=json.load(open('gpurun_out/bench_$P.json')); print({k:d[k] for k in ('value','ms_per_step','phases_ms','gpu_launches')}); print(d['e2e']); print(d['config'])
PY
tail -3 gpurun_out/bench_$P.err
done
bash tests/tools/gpu_ncu_list.sh 10000000 2>&1 | grep -v "^+"
