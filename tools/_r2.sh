mkdir -p gpurun_out
NCU="ncu --set full --import-source on --clock-control none"
$NCU -k regex:psf_fwd_tc_kernel --launch-skip 3 --launch-count 1 -o gpurun_out/r3_psf_fwd3 -f python tools/psf_probe.py 16384 > gpurun_out/r3_ncu_a.log 2>&1
$NCU -k regex:psf_fwd_tc_kernel --launch-skip 17 --launch-count 1 -o gpurun_out/r3_psf_fwd1 -f python tools/psf_probe.py 16384 > gpurun_out/r3_ncu_b.log 2>&1
$NCU -k regex:psf_bwd_tc_kernel --launch-skip 3 --launch-count 1 -o gpurun_out/r3_psf_bwd3 -f python tools/psf_probe.py 16384 > gpurun_out/r3_ncu_c.log 2>&1
tail -3 gpurun_out/r3_ncu_a.log gpurun_out/r3_ncu_c.log
ls -la gpurun_out/*.ncu-rep
