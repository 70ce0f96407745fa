mkdir -p gpurun_out
N=8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521"
echo "== bench N=8 overlapped"; timeout 300 $TR bench.py --gpus $N --steps 10 --warmup 3 --no-extras --no-cpu-baseline 2>gpurun_out/r2_n8_a.err | tee gpurun_out/r2_bench_8gpu_overlap.json | cut -c1-220
MS=$(python -c "import json; print(json.load(open('gpurun_out/r2_bench_8gpu_overlap.json'))['ms_per_step'])")
echo "== bench N=8 one all-reduce after backward"; TSR_DP_OVERLAP=0 timeout 300 $TR bench.py --gpus $N --steps 10 --warmup 3 --no-extras --no-cpu-baseline 2>gpurun_out/r2_n8_b.err | tee gpurun_out/r2_bench_8gpu_nooverlap.json | cut -c1-220
if python -c "import sys; sys.exit(0 if float('$MS') > 56.5 else 1)"; then
  echo "== bench N=8 overlapped, NCCL_MAX_CTAS=8"; NCCL_MAX_CTAS=8 timeout 300 $TR bench.py --gpus $N --steps 10 --warmup 3 --no-extras --no-cpu-baseline 2>gpurun_out/r2_n8_c.err | tee gpurun_out/r2_bench_8gpu_maxctas8.json | cut -c1-220
fi
echo "== C5 joint, 8192 per GPU x 8"; timeout 400 $TR tools/joint_c5_dp.py 8192 4 4 2>gpurun_out/r2_n8_c5.err | tee gpurun_out/r2_c5_8gpu.json | tail -2
echo "== C4 (S=7) strong scaling, global batch 2048, N=8"; timeout 300 $TR bench.py --gpus $N --steps 10 --warmup 3 --seqs 7 --global-batch 2048 --no-extras --no-cpu-baseline 2>gpurun_out/r2_n8_d.err | tee gpurun_out/r2_bench_c4_strong_8gpu.json | cut -c1-220
echo "== C4 (S=7) weak scaling, 1024 per GPU, N=8"; timeout 300 $TR bench.py --gpus $N --steps 10 --warmup 3 --seqs 7 --no-extras --no-cpu-baseline 2>gpurun_out/r2_n8_e.err | tee gpurun_out/r2_bench_c4_weak_8gpu.json | cut -c1-220
tail -n 3 gpurun_out/r2_n8_a.err gpurun_out/r2_n8_c5.err
