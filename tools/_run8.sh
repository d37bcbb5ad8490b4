# one call on a 2-GPU box: single-GPU validation / A-B runs on GPU 0, then the 2-GPU data-parallel checks
mkdir -p gpurun_out
L=gpurun_out/r2_c8.log
echo "=== nvidia-smi" > $L; nvidia-smi -L >> $L 2>&1
echo "=== attention sanity gate (persistent kernels); falls back to the previous build when it fails" >> $L
GATE=new
for mode in 3 1 0; do
  if DB200_ATTN_PERSIST=$mode timeout 240 python tools/gpu_diag.py attn > gpurun_out/attn_gate_$mode.log 2>&1 && grep -q " 0 bad" gpurun_out/attn_gate_$mode.log; then
    echo "attention OK with DB200_ATTN_PERSIST=$mode" >> $L; grep "PERF\|SUMMARY" gpurun_out/attn_gate_$mode.log >> $L
    export DB200_ATTN_PERSIST=$mode; break
  else
    echo "attention FAILED with DB200_ATTN_PERSIST=$mode" >> $L; grep "BAD\|SUMMARY\|rror" gpurun_out/attn_gate_$mode.log | head -20 >> $L
    if [ $mode = 0 ]; then GATE=prev; export DB200_LIB=$PWD/dalle_mtf_b200/libdalle_b200_prev.so; unset DB200_ATTN_PERSIST; echo "using libdalle_b200_prev.so" >> $L; fi
  fi
done
echo "=== kernel + engine tests (GEMM epilogue prefetch, fp32 tensor-core convs)" >> $L
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_engine_gpu.py -q -m gpu >> $L 2>&1; echo "exit=$?" >> $L
echo "=== bench A: this build (5 stages, cp.async operand prefetch)" >> $L
DB200_BENCH_VERBOSE=1 timeout 400 python bench.py --steps 10 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/bench_r02_c8_new.json 2> gpurun_out/bench_r02_c8_new.err; echo "exit=$?" >> $L
grep "gemm M=" gpurun_out/bench_r02_c8_new.err >> $L; grep "^{" gpurun_out/bench_r02_c8_new.json | cut -c1-400 >> $L
echo "=== bench B: previous GEMM configuration (6 stages, synchronous fetch)" >> $L
DB200_LIB=$PWD/dalle_mtf_b200/libdalle_b200_ab.so DB200_BENCH_VERBOSE=1 timeout 400 python bench.py --steps 10 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/bench_r02_c8_old.json 2> gpurun_out/bench_r02_c8_old.err; echo "exit=$?" >> $L
grep "gemm M=" gpurun_out/bench_r02_c8_old.err >> $L; grep "^{" gpurun_out/bench_r02_c8_old.json | cut -c1-400 >> $L
if [ $GATE = new ]; then
echo "=== attention variants (dev library)" >> $L
timeout 500 python tools/attn_experiments.py >> $L 2>&1; echo "exit=$?" >> $L
echo "=== attention timeline (dev library)" >> $L
DB200_LIB=$PWD/dalle_mtf_b200/libdalle_b200_dev.so timeout 200 python tools/attn_trace.py > gpurun_out/attn_trace_r02.txt 2>&1; echo "exit=$?" >> $L
DB200_LIB=$PWD/dalle_mtf_b200/libdalle_b200_dev.so DB200_ATTN_PERSIST=0 timeout 200 python tools/attn_trace.py > gpurun_out/attn_trace_r02_nonpersist.txt 2>&1; echo "exit=$?" >> $L
fi
echo "=== vae_example launch list" >> $L
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_vae_example_r02.csv python bench.py --vae-example --steps 2 --warmup 2 >> $L 2>&1; echo "exit=$?" >> $L
timeout 200 python bench.py --vae-example --steps 50 --warmup 10 >> $L 2>&1; echo "exit=$?" >> $L
echo "=== 2-GPU part" >> $L
bash tools/_run_mg2.sh > /dev/null 2>&1
cat gpurun_out/r2_mg2.log >> $L
tail -150 $L
