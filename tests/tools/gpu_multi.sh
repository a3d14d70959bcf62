# bench.py on N GPUs of one box under torchrun, as the driver launches it:  gpurun --gpus N -- 'bash tests/tools/gpu_multi.sh N [tag]'
N=${1:-2}; TAG=${2:-r2}
set -x
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_${N}gpu.json 2> gpurun_out/${TAG}_bench_${N}gpu.err
tail -3 gpurun_out/${TAG}_bench_${N}gpu.err
python - <<PY
import json
d=json.load(open('gpurun_out/${TAG}_bench_${N}gpu.json'))
print("weak: ms/step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["ms_per_step"], "parity", d["parity"])
o=d.get("one_stream") or {}
print("one stream: ms/step", o.get("ms_per_step"), o.get("parity"), o.get("exchange_rounds_per_step"))
for r in o.get("per_rank", []): print(r)
PY
