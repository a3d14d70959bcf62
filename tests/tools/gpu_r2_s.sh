set -x
export SQUID_CONFIG_TEST_PAIRS=300000
timeout 2400 compute-sanitizer --tool memcheck --print-limit 3 python -m pytest tests/test_gpu_parity.py tests/test_gpu_wire.py tests/test_gpu_sharded.py tests/test_gpu_configs.py tests/test_gpu_dropin.py tests/test_gpu_sort.py -q -m gpu --tb=line > gpurun_out/sanitizer_all.log 2>&1
grep "Invalid\|     at \|ERROR SUMMARY\|passed\|failed" gpurun_out/sanitizer_all.log | head -20
