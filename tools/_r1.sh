mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_tpsf_gpu.py tests/test_ops_gpu.py -x -q --timeout 300 -s 2>&1 | grep -v "^$" | tail -40 > gpurun_out/r3_psf_tests.log; tail -30 gpurun_out/r3_psf_tests.log
timeout 120 python tools/psf_probe.py 16384 2>&1 | tail -12 | tee gpurun_out/r3_psf_probe.log
