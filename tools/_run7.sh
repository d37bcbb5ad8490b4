mkdir -p gpurun_out
L=gpurun_out/r2_c7.log
echo "=== kernel tests (elect-issue pattern in attention, GEMM, conv)" > $L
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x >> $L 2>&1; echo "exit=$?" >> $L
echo "=== attention diag" >> $L
timeout 300 python tools/gpu_diag.py attn >> $L 2>&1; echo "exit=$?" >> $L
echo "=== attention timing (dev library)" >> $L
timeout 600 python tools/attn_experiments.py >> $L 2>&1; echo "exit=$?" >> $L
echo "=== fp32 conv on tcgen05 (3-way split) + engine tests" >> $L
timeout 300 python -m pytest tests/test_engine_gpu.py -q -m gpu -x -k "conv" >> $L 2>&1; echo "exit=$?" >> $L
timeout 600 python -m pytest tests/test_engine_gpu.py -q -m gpu >> $L 2>&1; echo "exit=$?" >> $L
echo "=== bench N=1" >> $L
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r02_n1b.json 2> gpurun_out/bench_r02_n1b.err; echo "exit=$?" >> $L
head -c 3000 gpurun_out/bench_r02_n1b.json >> $L
tail -150 $L
