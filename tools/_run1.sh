mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tc2_gpu.py -q -x --timeout 300 2>&1 | tail -40 > gpurun_out/r2_tc2_tests.log
timeout 300 python tools/tc2_perf.py 1024 > gpurun_out/r2_tc2_perf.log 2>&1
tail -5 gpurun_out/r2_tc2_tests.log; cat gpurun_out/r2_tc2_perf.log
