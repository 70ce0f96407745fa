mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 600 2>&1 | tail -40 > gpurun_out/r2_gputests3.log
tail -15 gpurun_out/r2_gputests3.log
timeout 200 python tools/tc2_perf.py 1024 > gpurun_out/r2_tc2_perf3.log 2>&1; cat gpurun_out/r2_tc2_perf3.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2_bench3.json 2> gpurun_out/r2_bench3.err; cat gpurun_out/r2_bench3.json; tail -5 gpurun_out/r2_bench3.err
