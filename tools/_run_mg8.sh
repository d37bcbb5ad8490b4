# 8-GPU measurements (charged 8x: keep it short): weak + strong dalle_example, the 12 B configuration, vae_coco
mkdir -p gpurun_out
L=gpurun_out/r2_mg8.log
echo "=== nvidia-smi" > $L; nvidia-smi -L >> $L 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521"
echo "=== data-parallel smoke (two communicators); falls back to one communicator, aborts if that fails too" >> $L
if timeout 150 $TR tools/dp_smoke.py > gpurun_out/dp_smoke8.log 2>&1 && grep -q "DP SMOKE OK" gpurun_out/dp_smoke8.log; then
  echo "smoke OK (tail communicator on)" >> $L
else
  echo "smoke FAILED with the tail communicator:" >> $L; tail -5 gpurun_out/dp_smoke8.log >> $L
  export DB200_NCCL_TAIL_CTAS=-1
  if timeout 150 $TR tools/dp_smoke.py > gpurun_out/dp_smoke8b.log 2>&1 && grep -q "DP SMOKE OK" gpurun_out/dp_smoke8b.log; then
    echo "smoke OK with DB200_NCCL_TAIL_CTAS=-1" >> $L
  else
    echo "smoke FAILED without it too: aborting the 8-GPU run" >> $L; tail -5 gpurun_out/dp_smoke8b.log >> $L; tail -20 $L; exit 1
  fi
fi
echo "=== bench N=8 weak" >> $L
NCCL_DEBUG=WARN timeout 240 $TR bench.py --gpus 8 --steps 10 --warmup 3 --no-extra > gpurun_out/bench_r02_n8_weak.json 2> gpurun_out/bench_r02_n8_weak.err; echo "exit=$?" >> $L
grep "^{" gpurun_out/bench_r02_n8_weak.json | cut -c1-900 >> $L; tail -3 gpurun_out/bench_r02_n8_weak.err >> $L
echo "=== bench N=8 strong" >> $L
timeout 200 $TR bench.py --gpus 8 --steps 10 --warmup 3 --no-extra --scaling strong > gpurun_out/bench_r02_n8_strong.json 2> gpurun_out/bench_r02_n8_strong.err; echo "exit=$?" >> $L
grep "^{" gpurun_out/bench_r02_n8_strong.json | cut -c1-900 >> $L; tail -3 gpurun_out/bench_r02_n8_strong.err >> $L
echo "=== bench N=8 dalle_12b (ZeRO-1)" >> $L
timeout 400 $TR bench.py --gpus 8 --steps 3 --warmup 3 --workload dalle_12b > gpurun_out/bench_r02_n8_12b.json 2> gpurun_out/bench_r02_n8_12b.err; echo "exit=$?" >> $L
grep "^{" gpurun_out/bench_r02_n8_12b.json | cut -c1-1200 >> $L; tail -5 gpurun_out/bench_r02_n8_12b.err >> $L
echo "=== bench N=8 vae_coco" >> $L
timeout 200 $TR bench.py --gpus 8 --steps 5 --warmup 3 --vae-coco > gpurun_out/bench_r02_n8_vae_coco.json 2> gpurun_out/bench_r02_n8_vae_coco.err; echo "exit=$?" >> $L
grep "^{" gpurun_out/bench_r02_n8_vae_coco.json | cut -c1-600 >> $L; tail -3 gpurun_out/bench_r02_n8_vae_coco.err >> $L
echo "=== bench N=8 dalle_coco" >> $L
timeout 200 $TR bench.py --gpus 8 --steps 5 --warmup 3 --workload dalle_coco --no-extra > gpurun_out/bench_r02_n8_dalle_coco.json 2> gpurun_out/bench_r02_n8_dalle_coco.err; echo "exit=$?" >> $L
grep "^{" gpurun_out/bench_r02_n8_dalle_coco.json | cut -c1-700 >> $L; tail -3 gpurun_out/bench_r02_n8_dalle_coco.err >> $L
tail -80 $L
