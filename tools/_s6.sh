mkdir -p gpurun_out
ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:conv2d_f32_kernel --launch-skip 8 --launch-count 1 -o gpurun_out/s6_conv_f32_tn8 -f python tools/step_profile.py 256 fp32 > gpurun_out/s6_ncu_a.log 2>&1
ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:conv2d_wgrad_f32_kernel --launch-skip 4 --launch-count 1 -o gpurun_out/s6_wgrad_f32 -f python tools/step_profile.py 256 fp32 > gpurun_out/s6_ncu_b.log 2>&1
tail -n 3 gpurun_out/s6_ncu_a.log gpurun_out/s6_ncu_b.log
