mkdir -p gpurun_out
L=gpurun_out/r2_c5.log
echo "=== attn diag NG=4 (default)" > $L
timeout 300 python tools/gpu_diag.py attn >> $L 2>&1; echo "exit=$?" >> $L
echo "=== attn diag NG=2" >> $L
DB200_ATTN_NG=2 timeout 300 python tools/gpu_diag.py attn >> $L 2>&1; echo "exit=$?" >> $L
echo "=== pytest attention" >> $L
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k attention -x >> $L 2>&1; echo "exit=$?" >> $L
echo "=== ncu full on attention kernels (NG=4)" >> $L
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_ -c 8 -o gpurun_out/prof_attn_r02c python tools/prof_attn.py >> $L 2>&1; echo "exit=$?" >> $L
echo "=== bench N=1 (no extras)" >> $L
DB200_BENCH_VERBOSE=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/bench_r02_n1b.json 2> gpurun_out/bench_r02_n1b.err; echo "exit=$?" >> $L
grep "^{" gpurun_out/bench_r02_n1b.json | cut -c1-400 >> $L; grep "gemm M" gpurun_out/bench_r02_n1b.err >> $L
grep -v "^\[OK \]" $L | grep -v "^==PROF" | tail -70
