mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tc2_gpu.py tests/test_sr_large_gpu.py tests/test_dp_gpu.py tests/test_sr_gpu.py -x -q --timeout 600 2>&1 | tail -30 > gpurun_out/r2_gputests5.log
tail -12 gpurun_out/r2_gputests5.log
timeout 300 python tools/tc2_perf.py 1024 > gpurun_out/r2_tc2_perf5.log 2>&1; cat gpurun_out/r2_tc2_perf5.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2_bench5.json 2> gpurun_out/r2_bench5.err; cut -c1-300 gpurun_out/r2_bench5.json; tail -3 gpurun_out/r2_bench5.err
