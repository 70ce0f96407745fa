mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tc2_gpu.py tests/test_sr_bf16_gpu.py tests/test_sr_gpu.py -x -q --timeout 600 2>&1 | tail -5 > gpurun_out/r2_gputests9.log; cat gpurun_out/r2_gputests9.log
timeout 300 python tools/tc2_perf.py 1024 2>&1 | grep -v CTA0 > gpurun_out/r2_tc2_perf9.log; cat gpurun_out/r2_tc2_perf9.log
timeout 300 python tools/graph_ab.py 1024 > gpurun_out/r2_graph_ab.log 2>&1; cat gpurun_out/r2_graph_ab.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2_bench9.json 2> gpurun_out/r2_bench9.err; python -c "
import json; d=json.load(open('gpurun_out/r2_bench9.json')); print(d['value'], d['ms_per_step'], d['step_breakdown_ms'])"; tail -3 gpurun_out/r2_bench9.err
