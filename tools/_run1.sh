mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2_c1_smi.txt 2>&1
for bits in 0 1 2 3; do
  echo "=== DB200_ATTN_V2=$bits" >> gpurun_out/r2_c1_attn.log
  DB200_ATTN_V2=$bits timeout 240 python tools/gpu_diag.py attn >> gpurun_out/r2_c1_attn.log 2>&1
  echo "exit=$?" >> gpurun_out/r2_c1_attn.log
done
tail -5 gpurun_out/r2_c1_attn.log
