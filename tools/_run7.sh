mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_sr_large_gpu.py tests/test_sr_bf16_gpu.py tests/test_loss_curve_gpu.py -q -s --timeout 900 -k "srcnn or overflow or loss_curve or c1 or b256" 2>&1 | grep -v "^$" | tail -40 > gpurun_out/r2_gputests7.log
tail -30 gpurun_out/r2_gputests7.log
timeout 120 python tools/tc2_case.py 1 wgrad5x5_128 && timeout 300 ncu --set full --clock-control none --import-source on -k regex:wgrad_tc3 -s 2 -c 1 -o gpurun_out/r2_wgrad5x5_128 python tools/tc2_case.py 1 wgrad5x5_128 > gpurun_out/ncu_w5.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:wgrad_tc3 -s 2 -c 1 -o gpurun_out/r2_wgrad5x5_64 python tools/tc2_case.py 1 wgrad5x5_64 > gpurun_out/ncu_w5b.log 2>&1
ls -la gpurun_out/r2_wgrad*
