mkdir -p gpurun_out
L=gpurun_out/r2_c4.log
echo "=== attn diag (ws v4, two-ring producers)" > $L
timeout 300 python tools/gpu_diag.py attn >> $L 2>&1; echo "exit=$?" >> $L
echo "=== ncu full on attention kernels" >> $L
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_ -c 8 -o gpurun_out/prof_attn_r02b python tools/prof_attn.py >> $L 2>&1; echo "exit=$?" >> $L
echo "=== pytest new / touched tests" >> $L
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_sampling_gpu.py tests/test_data_gpu.py -q -m gpu -x >> $L 2>&1; echo "exit=$?" >> $L
timeout 900 python -m pytest tests/test_engine_gpu.py -q -m gpu -x >> $L 2>&1; echo "exit=$?" >> $L
timeout 600 python -m pytest tests/test_parity_baseline_gpu.py -q -m gpu -x -s -k "12b or optimizer_options or quirk" >> $L 2>&1; echo "exit=$?" >> $L
echo "=== bench N=1" >> $L
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r02_n1.json 2> gpurun_out/bench_r02_n1.err; echo "exit=$?" >> $L
tail -c 3000 gpurun_out/bench_r02_n1.json >> $L; tail -5 gpurun_out/bench_r02_n1.err >> $L
echo "=== sampling bench" >> $L
timeout 300 python tools/bench_sampling.py >> $L 2>&1; echo "exit=$?" >> $L
grep -v "^\[OK \]" $L | grep -v "^==PROF" | tail -60
