mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/s12_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/s12_ncu.log 2>&1
tail -2 gpurun_out/s12_ncu.log | cut -c1-300
wc -l gpurun_out/s12_launches.csv
