set -x
python bench.py > gpurun_out/bench_final_default.json 2> gpurun_out/bench_final_default.err
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_final_reference.json 2> gpurun_out/bench_final_reference.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench_final.log 2>&1
python - <<'PY'
import csv,collections,json
d=json.load(open('gpurun_out/bench_final_default.json'))
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["roofline"]["kernel"], d["roofline"]["frac"], d["gpu_launches"], d["cpu_baseline"]["value"])
rows=[r for r in csv.reader(open('gpurun_out/launches_final.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); ui=hdr.index('Metric Unit')
agg=collections.OrderedDict()
for r in rows[1:]:
    v=float(r[vi].replace(',','')); u=r[ui]
    v = v/1e6 if u in ('ns','nsecond') else (v/1e3 if u in ('us','usecond') else v)
    k=r[ki][:60]; a=agg.setdefault(k,[0,0.0]); a[0]+=1; a[1]+=v
tot=sum(v for c,v in agg.values())
for k,(c,v) in sorted(agg.items(), key=lambda x:-x[1][1])[:22]: print("%9.3f ms %5.1f%% %4d  %s"%(v,100*v/tot,c,k))
PY
