mkdir -p gpurun_out
L=gpurun_out/r2_c4.log
echo "=== attn diag (ws v4)" > $L
timeout 300 python tools/gpu_diag.py attn >> $L 2>&1; echo "exit=$?" >> $L
echo "=== pytest attention" >> $L
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k attention -x >> $L 2>&1; echo "exit=$?" >> $L
echo "=== ncu full on attention kernels" >> $L
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_ -c 8 -o gpurun_out/prof_attn_r02b python tools/prof_attn.py >> $L 2>&1; echo "exit=$?" >> $L
grep -v "^\[OK \]" $L | grep -v "^==PROF" | tail -40
