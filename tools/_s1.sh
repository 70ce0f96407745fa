mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q --timeout 600 --durations=8 2>&1 | tail -25 ) > gpurun_out/s1_gputests.log 2>&1
tail -5 gpurun_out/s1_gputests.log
timeout 120 python tools/psf_probe.py 16384 2>&1 | tee gpurun_out/s1_psf_probe.log
( time timeout 600 python bench.py ) > gpurun_out/s1_bench.log 2>&1; tail -c 3000 gpurun_out/s1_bench.log
