mkdir -p gpurun_out
for T in 256 512; do
echo "== threads $T"
TSR_PSF_THREADS=$T timeout 600 python -m pytest tests/test_tpsf_gpu.py tests/test_ops_gpu.py -x -q --timeout 300 2>&1 | tail -4
TSR_PSF_THREADS=$T timeout 120 python tools/psf_probe.py 16384 2>&1 | grep -v ffma | tee gpurun_out/r3_psf_probe3_$T.log
done
