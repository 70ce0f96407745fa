mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_sr_large_gpu.py tests/test_loss_curve_gpu.py tests/test_sr_bf16_gpu.py -q -s --timeout 900 -k "srcnn or loss_curve or overflow or c1 or b256" > gpurun_out/r2_gputests8.log 2>&1
grep -E "smoothed|worst|TactileSRCNN|passed|failed" gpurun_out/r2_gputests8.log | cut -c1-300
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench8.json 2> gpurun_out/r2_bench8.err; cat gpurun_out/r2_bench8.json; tail -3 gpurun_out/r2_bench8.err
