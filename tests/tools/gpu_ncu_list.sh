set -x
P=${1:-10000000}
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'^k_|Device' -c 300 --csv --log-file gpurun_out/launches.csv python bench.py --pairs $P --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
python - <<'PY'
import csv,collections
rows=[r for r in csv.reader(open('gpurun_out/launches.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); ui=hdr.index('Metric Unit')
agg=collections.OrderedDict()
seq=[]
for r in rows[1:]:
    v=float(r[vi].replace(',','')); u=r[ui]
    v = v/1e6 if u in ('ns','nsecond') else (v/1e3 if u in ('us','usecond') else v)
    k=r[ki][:70]; a=agg.setdefault(k,[0,0.0]); a[0]+=1; a[1]+=v; seq.append((k,v))
print("unit sample:", rows[1][ui])
for k,(c,v) in sorted(agg.items(), key=lambda x:-x[1][1])[:25]: print("%10.3f ms %4d  %s"%(v,c,k))
PY
