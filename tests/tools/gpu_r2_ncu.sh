# round-2 profiles: bench line, ncu launch list of the same command, ncu --set full of the stream kernels
set -x
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_g_bench.json 2> gpurun_out/r2_g_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_g_bench.json'))
print("ms/step", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], "phases", {k: round(v,2) for k,v in d["phases_ms"].items()}, "timeline", {k: round(v,2) for k,v in d["host_timeline_ms"].items()}, d["stats"]["raw_edges"], d["parity"]["ok"])
PY
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:'sq::|cub::|^k_|gsort' -c 1600 --csv --log-file gpurun_out/r2_launches_100M.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
python - <<'PY'
import csv,collections
rows=[r for r in csv.reader(open('gpurun_out/r2_launches_100M.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); ui=hdr.index('Metric Unit')
seq=[]
for r in rows[1:]:
    v=float(r[vi].replace(',','')); u=r[ui]
    v = v/1e6 if u in ('ns','nsecond') else (v/1e3 if u in ('us','usecond') else v)
    seq.append((r[ki][:90],v))
idx=[i for i,(k,v) in enumerate(seq) if 'k_classify_tiles' in k]
print("classify launches at", idx)
a,b=(idx[-2],idx[-1]) if len(idx)>=2 else (0,len(seq))
agg=collections.OrderedDict()
for k,v in seq[a:b]:
    if 'gsort' in k: k='gsort::*'
    x=agg.setdefault(k,[0,0.0]); x[0]+=1; x[1]+=v
tot=sum(v for c,v in agg.values())
print("one step: %d launches, %.2f ms of kernels" % (b-a, tot))
for k,(c,v) in sorted(agg.items(), key=lambda x:-x[1][1])[:28]: print("%8.3f ms %5.1f%% %4d %s"%(v,100*v/tot,c,k))
PY
timeout 1500 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'k_classify_tiles|k_assign_tiles|k_edges_generic|k_seed_islands|k_cov_compact|k_cov_count_tiles|k_rest_collect' -c 9 -o gpurun_out/r2_prof_20M python bench.py --pairs 20000000 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out/*.ncu-rep
