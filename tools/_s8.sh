mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_sr_gpu.py -x -q --timeout 300 2>&1 | tail -4 | tee gpurun_out/s8_tests.log
ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:conv2d_f32_kernel --launch-skip 14 --launch-count 3 -o gpurun_out/s8_conv_f32 -f python tools/step_profile.py 256 fp32 > gpurun_out/s8_ncu_a.log 2>&1
ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:conv2d_wgrad_f32_kernel --launch-skip 4 --launch-count 2 -o gpurun_out/s8_wgrad_f32 -f python tools/step_profile.py 256 fp32 > gpurun_out/s8_ncu_b.log 2>&1
tail -n 2 gpurun_out/s8_ncu_a.log gpurun_out/s8_ncu_b.log
