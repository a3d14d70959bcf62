set -x
python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3
SQG_SLOW_IN_TILE=1 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "seeded or golden or short" 2>&1 | tail -3
SQG_TIMING=1 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r2_h_bench.json 2> gpurun_out/r2_h_bench.err
grep '\[sqg\]' gpurun_out/r2_h_bench.err | tail -12
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_h_bench.json'))
print("ms/step", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], "phases", {k: round(v,2) for k,v in d["phases_ms"].items()}, "timeline", {k: round(v,2) for k,v in d["host_timeline_ms"].items()}, d["stats"]["raw_edges"], d["parity"]["ok"])
PY
SQUID_NVCC_EXTRA="-DSQ_SEED_PROF" python -m squid_b200.build --force > gpurun_out/r2_probe_build.log 2>&1
SQG_SEED_PROF_OUT=gpurun_out/seed_prof.bin python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2_probe_bench_prof.json 2> gpurun_out/r2_probe_bench_prof.err
python - <<'PY'
import numpy as np
raw = open('gpurun_out/seed_prof.bin','rb').read()
n = int(np.frombuffer(raw[:4], dtype=np.int32)[0])
span = np.frombuffer(raw[4:4+4*n], dtype=np.int32)
q = np.frombuffer(raw[4+4*n:4+4*n+8*12*n], dtype=np.int64).reshape(n, 12)
dur = (q[:,1]-q[:,0]) / 1e6
t0 = q[:,0][q[:,0]>0].min()
order = np.argsort(-dur)[:14]
names = ["replay","setup","margins","sort","tabulate","breakloop","consume","extend"]
print("islands", n, "total island-ms", dur.sum(), "kernel span ms", (q[:,1].max()-t0)/1e6)
for i in order:
    cyc = q[i,4:12].astype(float); tot = cyc.sum() or 1
    print("isl %6d span %8d groups %4d W %4d start %.3f dur %.3f ms | " % (i, span[i], q[i,2], q[i,3], (q[i,0]-t0)/1e6, dur[i]) + " ".join("%s %.0f%%" % (nm, 100*c/tot) for nm,c in zip(names,cyc)))
for W in (32, 512):
    m = q[:,3]==W
    if m.any(): print("W", W, "count", m.sum(), "sum ms", dur[m].sum(), "max ms", dur[m].max(), "last end ms", (q[m,1].max()-t0)/1e6)
PY
