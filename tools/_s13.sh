mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sr_bf16_gpu.py tests/test_graph_gpu.py tests/test_sr_large_gpu.py -x -q --timeout 300 2>&1 | tail -3 | tee gpurun_out/s13_tests.log
( timeout 300 python bench.py --no-extras --no-cpu-baseline ) > gpurun_out/s13_bench.log 2>&1; grep -o '"step_breakdown_ms.*' gpurun_out/s13_bench.log | cut -c1-500; grep -o '"value": [0-9.]*' gpurun_out/s13_bench.log | head -1
