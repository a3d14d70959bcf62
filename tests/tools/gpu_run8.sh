set -x
python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 900 python bench.py --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/bench_100M.json 2> gpurun_out/bench_100M.err; python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_100M.json')); print({k:d[k] for k in ('value','ms_per_step','phases_ms','gpu_launches','stats')}); print(d['e2e']); print(d['roofline'])
except Exception as e: print("bench failed", e)
PY
tail -3 gpurun_out/bench_100M.err
for K in k_classify k_conc_edges k_depth_targets k_cov_count; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -c 1 -o gpurun_out/prof_$K -f python bench.py --pairs 20000000 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_$K.log 2>&1
ncu -i gpurun_out/prof_$K.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin))
if len(rows)>2:
    h=rows[0]; v=rows[2] if len(rows)>2 else rows[1]
    want=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','l1tex__t_sector_hit_rate.pct','lts__t_sector_hit_rate.pct','sm__throughput.avg.pct_of_peak_sustained_elapsed','launch__occupancy_limit_registers','smsp__cycles_active.avg']
    for w in want:
        for i,c in enumerate(h):
            if c==w: print('$K',w,rows[1][i],v[i])
"
done
ls -la gpurun_out/*.ncu-rep
