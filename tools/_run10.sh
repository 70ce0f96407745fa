mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
echo "== bench N=2 overlapped"; timeout 300 $TR bench.py --gpus 2 --steps 5 --warmup 3 --no-extras --no-cpu-baseline 2>gpurun_out/r2_n2_a.err | cut -c1-200
echo "== bench N=2 one all-reduce"; TSR_DP_OVERLAP=0 timeout 300 $TR bench.py --gpus 2 --steps 5 --warmup 3 --no-extras --no-cpu-baseline 2>gpurun_out/r2_n2_b.err | cut -c1-200
echo "== bench N=2 B=32 graph"; timeout 300 $TR bench.py --gpus 2 --steps 20 --warmup 5 --batch 32 --cuda-graph --no-extras --no-cpu-baseline 2>gpurun_out/r2_n2_c.err | cut -c1-200
echo "== bench N=2 B=32 eager"; timeout 300 $TR bench.py --gpus 2 --steps 20 --warmup 5 --batch 32 --no-extras --no-cpu-baseline 2>gpurun_out/r2_n2_d.err | cut -c1-200
echo "== entry point SR DP + graph"; timeout 300 $TR -m tactilesr_b200.train.tactileSR_train --synthetic 512 --epochs 2 --train_batch_size 32 --save_dir gpurun_out/tmp_sr --cuda-graph 2>&1 | tail -6
echo "== entry point tPSFNet"; timeout 300 python -m tactilesr_b200.train.tPSFNet_train --synthetic 1024 --epochs 2 --save_dir gpurun_out/tmp_psf 2>&1 | tail -4
echo "== entry point Seqs"; timeout 300 python -m tactilesr_b200.train.tactileSRSeqs_train --synthetic 128 --epochs 1 --train_batch_size 32 --save_dir gpurun_out/tmp_seqs 2>&1 | tail -4
echo "== C5 DP 2 GPUs"; timeout 300 $TR tools/joint_c5_dp.py 2048 2 3 2>&1 | tail -3
tail -3 gpurun_out/r2_n2_a.err gpurun_out/r2_n2_c.err
rm -rf gpurun_out/tmp_sr gpurun_out/tmp_psf gpurun_out/tmp_seqs
