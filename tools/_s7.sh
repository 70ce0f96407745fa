mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_tpsf_gpu.py tests/test_ops_gpu.py -x -q --timeout 300 2>&1 | tail -4 | tee gpurun_out/s7_tests.log
timeout 120 python tools/psf_probe.py 16384 2>&1 | grep -v ffma | tee gpurun_out/s7_psf_probe.log
( time timeout 600 python bench.py ) > gpurun_out/s7_bench.log 2>&1; tail -c 2500 gpurun_out/s7_bench.log
