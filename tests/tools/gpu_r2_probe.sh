# round-2 probe: host laps of one step, seed-machine per-island profile (SQ_SEED_PROF build), at the bench size
set -x
P=${1:-100000000}
SQG_TIMING=1 python bench.py --pairs $P --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/r2_probe_bench.json 2> gpurun_out/r2_probe_bench.err
tail -c 3000 gpurun_out/r2_probe_bench.json
grep '\[sqg\]' gpurun_out/r2_probe_bench.err | tail -60
# profile build
SQUID_NVCC_EXTRA="-DSQ_SEED_PROF" python -m squid_b200.build --force > gpurun_out/r2_probe_build.log 2>&1
SQG_SEED_PROF_OUT=gpurun_out/seed_prof.bin python bench.py --pairs $P --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2_probe_bench_prof.json 2> gpurun_out/r2_probe_bench_prof.err
python - <<'PY'
import numpy as np
raw = open('gpurun_out/seed_prof.bin','rb').read()
n = int(np.frombuffer(raw[:4], dtype=np.int32)[0])
span = np.frombuffer(raw[4:4+4*n], dtype=np.int32)
q = np.frombuffer(raw[4+4*n:4+4*n+8*12*n], dtype=np.int64).reshape(n, 12)
dur = (q[:,1]-q[:,0]) / 1e6
t0 = q[:,0][q[:,0]>0].min()
order = np.argsort(-dur)[:25]
names = ["replay","setup","margins","sort","tabulate","breakloop","consume","extend"]
print("islands", n, "total island-ms", dur.sum(), "kernel span ms", (q[:,1].max()-t0)/1e6)
for i in order:
    cyc = q[i,4:12].astype(float); tot = cyc.sum() or 1
    print("isl %6d span %8d groups %4d W %4d start %.3f dur %.3f ms | " % (i, span[i], q[i,2], q[i,3], (q[i,0]-t0)/1e6, dur[i]) + " ".join("%s %.0f%%" % (nm, 100*c/tot) for nm,c in zip(names,cyc)))
# histogram of durations by W
for W in (32, 512):
    m = q[:,3]==W
    if m.any(): print("W", W, "count", m.sum(), "sum ms", dur[m].sum(), "max ms", dur[m].max())
PY
