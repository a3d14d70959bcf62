set -x
timeout 1500 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'k_classify_tiles|k_assign_tiles|k_edges_generic|k_cov_count_tiles' -c 5 -o gpurun_out/r2_prof2_20M python bench.py --pairs 20000000 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_full2.log 2>&1
ls -la gpurun_out/*.ncu-rep
