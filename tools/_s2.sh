mkdir -p gpurun_out
NCU="ncu --set full --import-source on --clock-control none"
# probe order: f_tc x7 launches (2 warm + 5), f_tca x7, f_16 x7, f_16a x7, f_ff x7, b_tc x7, b_16 x7
$NCU -k regex:psf_fwd_tc_kernel --launch-skip 3 --launch-count 1 -o gpurun_out/s2_psf_fwd3 -f python tools/psf_probe.py 16384 > gpurun_out/s2_ncu_a.log 2>&1
$NCU -k regex:psf_bwd_tc_kernel --launch-skip 3 --launch-count 1 -o gpurun_out/s2_psf_bwd3 -f python tools/psf_probe.py 16384 > gpurun_out/s2_ncu_c.log 2>&1
tail -3 gpurun_out/s2_ncu_a.log gpurun_out/s2_ncu_c.log
ls -la gpurun_out/*.ncu-rep
