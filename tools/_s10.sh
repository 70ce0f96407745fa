mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q --timeout 600 2>&1 | tail -6 ) > gpurun_out/s10_gputests.log 2>&1
tail -4 gpurun_out/s10_gputests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -8 | tee gpurun_out/s10_smoke.log
( time timeout 600 python bench.py ) > gpurun_out/s10_bench.log 2>&1; tail -c 1200 gpurun_out/s10_bench.log
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -2 | cut -c1-600
