mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sr_gpu.py tests/test_sr_bf16_gpu.py tests/test_graph_gpu.py tests/test_sr_large_gpu.py tests/test_dp_gpu.py -x -q --timeout 300 2>&1 | tail -4 | tee gpurun_out/s9_tests.log
timeout 100 python tools/headtail_perf.py 1024 2>&1 | tee gpurun_out/s9_headtail.log
( timeout 300 python bench.py --no-extras --no-cpu-baseline ) > gpurun_out/s9_bench.log 2>&1; tail -c 1500 gpurun_out/s9_bench.log
