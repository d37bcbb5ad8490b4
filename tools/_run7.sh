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
echo "=== ncu full: tokenizer convs + first GEMMs of a step" >> $L
timeout 420 ncu --set full --clock-control none --import-source on -k regex:'conv_tc|conv_first' -c 15 -o gpurun_out/prof_convtc_r02 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extra >> $L 2>&1; echo "exit=$?" >> $L
timeout 420 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 78 -c 14 -o gpurun_out/prof_gemm_r02 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extra >> $L 2>&1; echo "exit=$?" >> $L
tail -12 $L
