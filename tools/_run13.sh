mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_tc2_gpu.py tests/test_sr_bf16_gpu.py -x -q --timeout 600 2>&1 | tail -3
timeout 300 python tools/tc2_perf.py 1024 2>&1 | grep -v CTA0 > gpurun_out/r2_tc2_perf13.log; cat gpurun_out/r2_tc2_perf13.log
for v in A B C; do
  if [ $v = A ]; then export TSR_BENCH_NO_CLOCKS=0; fi
  if [ $v = B ]; then export TSR_BENCH_NO_CLOCKS=1; fi
  if [ $v = C ]; then export TSR_BENCH_NO_CLOCKS=0; export TSR_BENCH_CLOCK_MS=1000; fi
  timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/r2_bench13$v.json 2> gpurun_out/r2_bench13$v.err; python -c "
import json; d=json.load(open('gpurun_out/r2_bench13$v.json')); print('$v', d['value'], d['ms_per_step'], d['clocks'], d['step_breakdown_ms'])"
done
timeout 200 python tools/graph_ab.py 1024
