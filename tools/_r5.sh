mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_tpsf_gpu.py tests/test_ops_gpu.py -x -q --timeout 300 2>&1 | tail -4
timeout 120 python tools/psf_probe.py 16384 2>&1 | grep -v ffma | tee gpurun_out/r3_psf_probe4.log
