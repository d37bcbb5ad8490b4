# final evidence run on one GPU: full GPU test suite, smoke(), bench lines, ncu per-kernel metrics of one step
mkdir -p gpurun_out
L=gpurun_out/r2_final.log
echo "=== pytest -m gpu" > $L
timeout 1200 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu_r02.log 2>&1; echo "exit=$?" >> $L; tail -4 gpurun_out/pytest_gpu_r02.log >> $L
echo "=== smoke()" >> $L
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" >> $L 2>&1; echo "exit=$?" >> $L
echo "=== bench (default)" >> $L
timeout 900 python bench.py > gpurun_out/bench_r02_n1.json 2> gpurun_out/bench_r02_n1.err; echo "exit=$?" >> $L
grep "^{" gpurun_out/bench_r02_n1.json | cut -c1-400 >> $L
echo "=== bench --impl reference" >> $L
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r02_ref.json 2> gpurun_out/bench_r02_ref.err; echo "exit=$?" >> $L
grep "^{" gpurun_out/bench_r02_ref.json | cut -c1-400 >> $L
echo "=== sampling bench" >> $L
timeout 300 python tools/bench_sampling.py > gpurun_out/bench_r02_sampling.json 2>> $L; echo "exit=$?" >> $L; cat gpurun_out/bench_r02_sampling.json >> $L
echo "=== ncu: per-kernel metrics of one dalle_example step" >> $L
python -c "import bench; print(bench.csrc_hash())" > gpurun_out/csrc_hash.txt 2>> $L
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed
timeout 900 ncu --metrics $M --clock-control none -c 1400 --csv --log-file gpurun_out/ncu_step_r02.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extra >> $L 2>&1; echo "exit=$?" >> $L
echo "=== ncu: per-kernel metrics of one vae_coco step" >> $L
timeout 600 ncu --metrics $M --clock-control none -c 700 --csv --log-file gpurun_out/ncu_vaecoco_r02.csv \
  python bench.py --vae-coco --steps 1 --warmup 1 >> $L 2>&1; echo "exit=$?" >> $L
tail -60 $L
