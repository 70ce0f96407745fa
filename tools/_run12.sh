mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -8 > gpurun_out/r2_gputests12.log; cat gpurun_out/r2_gputests12.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench12.json 2> gpurun_out/r2_bench12.err; python -c "
import json; d=json.load(open('gpurun_out/r2_bench12.json')); print(d['value'], d['ms_per_step'], d['tensor_roofline_frac_step'], d['roofline']['frac'], d['e2e']['value']); print(d['step_breakdown_ms'])"; tail -3 gpurun_out/r2_bench12.err
timeout 300 python bench.py --steps 10 --warmup 3 --seqs 7 --global-batch 2048 --no-extras --no-cpu-baseline > gpurun_out/r2_bench_c4_strong_1gpu.json 2>/dev/null; cut -c1-200 gpurun_out/r2_bench_c4_strong_1gpu.json
timeout 300 python bench.py --steps 10 --warmup 3 --seqs 7 --no-extras --no-cpu-baseline > gpurun_out/r2_bench_c4_weak_1gpu.json 2>/dev/null; cut -c1-200 gpurun_out/r2_bench_c4_weak_1gpu.json
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2_plain12.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches12.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2_ncu12.log 2>&1
timeout 120 python tools/tc2_case.py 2 fwd5x5_128 && timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_tc2 -s 2 -c 1 -o gpurun_out/r2_conv_tc2_5x5_128_b1024 python tools/tc2_case.py 2 fwd5x5_128 > gpurun_out/ncu_f5.log 2>&1
ls -la gpurun_out/r2_launches12.csv gpurun_out/r2_conv_tc2_5x5_128_b1024.ncu-rep
