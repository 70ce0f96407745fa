mkdir -p gpurun_out
ncu --set full --import-source on --clock-control none -k regex:tail_fwd_blocked_kernel --launch-skip 5 --launch-count 1 -o gpurun_out/s15_tail_fwd -f python tools/headtail_perf.py 1024 > gpurun_out/s15_ncu.log 2>&1
tail -n 2 gpurun_out/s15_ncu.log
