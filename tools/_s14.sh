mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_headtail_gpu.py tests/test_sr_gpu.py tests/test_sr_bf16_gpu.py -x -q --timeout 300 2>&1 | tail -12 | tee gpurun_out/s14_tests.log
timeout 100 python tools/headtail_perf.py 1024 2>&1 | tee gpurun_out/s14_headtail.log
( timeout 300 python bench.py --no-extras --no-cpu-baseline ) > gpurun_out/s14_bench.log 2>&1; grep -o '"step_breakdown_ms.*' gpurun_out/s14_bench.log | cut -c1-500; grep -o '"value": [0-9.]*' gpurun_out/s14_bench.log | head -1
