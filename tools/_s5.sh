mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_sr_gpu.py tests/test_ops_gpu.py tests/test_tpsf_gpu.py tests/test_graph_gpu.py tests/test_sr_large_gpu.py -x -q --timeout 300 2>&1 | tail -4 | tee gpurun_out/s5_tests.log
timeout 200 python tools/fp32_mode_perf.py 2>&1 | tee gpurun_out/s5_fp32_perf.log
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/s5_launches_fp32_b256.csv python tools/step_profile.py 256 fp32 > gpurun_out/s5_ncu.log 2>&1
python tools/launch_summary.py gpurun_out/s5_launches_fp32_b256.csv | tee gpurun_out/s5_launches_fp32_b256.txt | head -8
