mkdir -p gpurun_out
L=gpurun_out/r2_c2.log
echo "=== attn diag (ws fwd)" > $L
timeout 300 python tools/gpu_diag.py attn >> $L 2>&1; echo "exit=$?" >> $L
echo "=== pytest attention" >> $L
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k attention -x >> $L 2>&1; echo "exit=$?" >> $L
echo "=== pytest parity baseline" >> $L
timeout 900 python -m pytest tests/test_parity_baseline_gpu.py -q -m gpu -s >> $L 2>&1; echo "exit=$?" >> $L
tail -30 $L
