set -x
python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -5
timeout 600 python bench.py --pairs 10000000 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_10M.json 2> gpurun_out/bench_10M.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_10M.json')); print({k:d[k] for k in ('value','ms_per_step','phases_ms','gpu_launches')}); print(d['e2e']); print(d['config'])
PY
tail -3 gpurun_out/bench_10M.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_10M.csv python bench.py --pairs 10000000 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
python - <<'PY'
import csv,collections
rows=[r for r in csv.reader(open('gpurun_out/launches_10M.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); ui=hdr.index('Metric Unit')
agg=collections.OrderedDict()
for r in rows[1:]:
    v=float(r[vi].replace(',','')); u=r[ui]
    v = v/1e6 if u=='ns' else (v/1e3 if u=='us' else v)
    k=r[ki][:90]; a=agg.setdefault(k,[0,0.0]); a[0]+=1; a[1]+=v
for k,(c,v) in sorted(agg.items(), key=lambda x:-x[1][1])[:30]: print("%9.3f ms %4d  %s"%(v,c,k))
PY
