set -x
python __graft_entry__.py smoke 2>&1 | tail -5
for P in 2000000 10000000; do
  timeout 900 python bench.py --pairs $P --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_$P.json 2> gpurun_out/bench_$P.err; tail -c 3000 gpurun_out/bench_$P.json; tail -5 gpurun_out/bench_$P.err
done
