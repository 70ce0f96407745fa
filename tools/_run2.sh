mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_tc2_gpu.py -q -x --timeout 300 -k wgrad 2>&1 | tail -5 > gpurun_out/r2_tc2_tests2.log
for na in 3 4; do echo "NA=$na"; TSR_TC2_NA=$na timeout 120 python tools/tc2_perf.py 1024 fwd; done > gpurun_out/r2_tc2_perf_na.log 2>&1
for g in 2 1; do
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 2 -c 1 -o gpurun_out/r2_dgrad1x1_gen$g python tools/tc2_case.py $g dgrad1x1 > gpurun_out/ncu_gen$g.log 2>&1
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_tc2 -s 2 -c 1 -o gpurun_out/r2_fwd3x3_128 python tools/tc2_case.py 2 fwd3x3_128 > gpurun_out/ncu_f3.log 2>&1
cat gpurun_out/r2_tc2_tests2.log gpurun_out/r2_tc2_perf_na.log; ls -la gpurun_out/*.ncu-rep
