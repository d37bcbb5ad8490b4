mkdir -p gpurun_out
L=gpurun_out/r2_c3.log
echo "=== attn diag (ws fwd + ws bwd)" > $L
timeout 300 python tools/gpu_diag.py attn >> $L 2>&1; echo "exit=$?" >> $L
echo "=== ncu full on attention kernels" >> $L
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_ -s 8 -c 8 -o gpurun_out/prof_attn_r02a python tools/prof_attn.py >> $L 2>&1; echo "exit=$?" >> $L
grep -v "^\[OK \]" $L | tail -40
