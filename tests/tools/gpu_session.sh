# One measurement session on a B200 box (run through gpurun from the repo root): GPU tests, the bench line of both arms, the ncu
# launch list of the bench command and one `ncu --set full` capture of the stream kernels.  Everything lands in gpurun_out/;
# the ncu captures are summarised on the box (tests/tools/ncu_summary.py) and only their CSV pages come back -- gpurun merges at most
# 64 MiB.  The copy into profiles/ is done afterwards, where the numbers are read.
#   bash tests/tools/gpu_session.sh [tag]        (tag defaults to r2)
TAG=${1:-r2}
set -x
timeout 2400 python -m pytest tests -x -q -m gpu > gpurun_out/${TAG}_pytest_gpu.log 2>&1; tail -3 gpurun_out/${TAG}_pytest_gpu.log
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -2 gpurun_out/${TAG}_bench.err
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:'sq::|cub::|^k_|gsort' -c 1800 --csv --log-file gpurun_out/${TAG}_launches_100M.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu_launches.log 2>&1
timeout 1500 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'k_classify_tiles|k_assign_tiles|k_edges_generic|k_seed_islands|k_cov_gather|k_cov_count_tiles' -c 11 -o gpurun_out/${TAG}_prof_20M python bench.py --pairs 20000000 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/${TAG}_ncu_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'k_wire_decode' -c 2 -o gpurun_out/${TAG}_prof_wire_20M python bench.py --pairs 20000000 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/${TAG}_ncu_wire.log 2>&1
REC=$(grep -o '"records": [0-9]*' gpurun_out/${TAG}_ncu_full.log | head -1 | grep -o '[0-9]*$')
python tests/tools/ncu_summary.py gpurun_out/${TAG}_prof_20M.ncu-rep --records $REC --traffic-json gpurun_out/${TAG}_traffic.json > gpurun_out/${TAG}_ncu_summary.txt
python tests/tools/ncu_summary.py gpurun_out/${TAG}_prof_wire_20M.ncu-rep --records $REC > gpurun_out/${TAG}_ncu_wire_decode.txt
ncu -i gpurun_out/${TAG}_prof_20M.ncu-rep --page raw --csv > gpurun_out/${TAG}_ncu_full_stream_kernels_20Mpairs.csv 2>/dev/null
rm -f gpurun_out/*.ncu-rep
ls -la gpurun_out/${TAG}_*
python - <<PY
import json
d=json.load(open('gpurun_out/${TAG}_bench.json'))
print("ms/step", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], "soa", d["e2e_soa"]["ms_per_step"], "parity", d["parity"], "clocks", d["clocks"])
print("phases", {k: round(v,2) for k,v in d["phases_ms"].items()})
print("roof", d["roofline"]["kernel"], d["roofline"]["frac"], d["roofline"]["whole_path"]["frac"])
print("cpu", json.dumps(d["cpu_baseline"])[:900])
print(open('gpurun_out/${TAG}_bench_reference.json').read()[:600])
PY
