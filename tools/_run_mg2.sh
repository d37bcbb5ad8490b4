# 2-GPU validation: hand-driven NCCL communicator, DP equivalence, ZeRO-1, weak / strong bench lines
mkdir -p gpurun_out
L=gpurun_out/r2_mg2.log
echo "=== nvidia-smi" > $L; nvidia-smi -L >> $L 2>&1
echo "=== pytest 2-GPU" >> $L
timeout 360 python -m pytest tests/test_dp_gpu.py -q -m gpu -x >> $L 2>&1; echo "exit=$?" >> $L
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
echo "=== bench N=2 weak" >> $L
NCCL_DEBUG=WARN timeout 300 $TR bench.py --gpus 2 --steps 10 --warmup 3 --no-extra > gpurun_out/bench_r02_n2_weak.json 2> gpurun_out/bench_r02_n2_weak.err; echo "exit=$?" >> $L
grep "^{" gpurun_out/bench_r02_n2_weak.json | cut -c1-700 >> $L; tail -3 gpurun_out/bench_r02_n2_weak.err >> $L
echo "=== bench N=2 strong" >> $L
timeout 300 $TR bench.py --gpus 2 --steps 10 --warmup 3 --no-extra --scaling strong > gpurun_out/bench_r02_n2_strong.json 2> gpurun_out/bench_r02_n2_strong.err; echo "exit=$?" >> $L
grep "^{" gpurun_out/bench_r02_n2_strong.json | cut -c1-700 >> $L; tail -3 gpurun_out/bench_r02_n2_strong.err >> $L
echo "=== bench N=2 dalle_12b width, 4 layers (ZeRO-1 smoke)" >> $L
timeout 480 $TR bench.py --gpus 2 --steps 3 --warmup 3 --workload dalle_12b --n-layers 4 > gpurun_out/bench_r02_n2_12b_l4.json 2> gpurun_out/bench_r02_n2_12b_l4.err; echo "exit=$?" >> $L
grep "^{" gpurun_out/bench_r02_n2_12b_l4.json | cut -c1-900 >> $L; tail -5 gpurun_out/bench_r02_n2_12b_l4.err >> $L
echo "=== bench N=2 vae_coco" >> $L
timeout 300 $TR bench.py --gpus 2 --steps 5 --warmup 3 --vae-coco > gpurun_out/bench_r02_n2_vae_coco.json 2> gpurun_out/bench_r02_n2_vae_coco.err; echo "exit=$?" >> $L
grep "^{" gpurun_out/bench_r02_n2_vae_coco.json | cut -c1-500 >> $L; tail -3 gpurun_out/bench_r02_n2_vae_coco.err >> $L
tail -60 $L
