# Evidence capture for profiles/ (run once on the final sources): launch list + per-kernel DRAM/tensor-pipe metrics of
# one whole DALL-E step and one tokenizer+VAE pass, and `--set full` captures of the attention kernels.
mkdir -p gpurun_out
L=gpurun_out/r2_ncu.log
python -c "import bench; print(bench.csrc_hash())" > gpurun_out/csrc_hash.txt 2>> $L
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed
echo "=== step metrics (dalle_example, 1 warm-up + 1 timed step)" > $L
timeout 900 ncu --metrics $M --clock-control none -c 1200 --csv --log-file gpurun_out/ncu_step_r02.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extra >> $L 2>&1; echo "exit=$?" >> $L
echo "=== step metrics (vae_coco: bf16 conv_tc / wgrad_tc kernels)" >> $L
timeout 900 ncu --metrics $M --clock-control none -k regex:'conv|gumbel|mse|rowmatmul|colsum|adam|space' -c 600 --csv --log-file gpurun_out/ncu_vaecoco_r02.csv \
  python bench.py --vae-coco --steps 1 --warmup 1 --no-cpu-baseline --no-extra >> $L 2>&1; echo "exit=$?" >> $L
echo "=== ncu full on attention kernels" >> $L
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_ -c 8 -o gpurun_out/prof_attn_r02f python tools/prof_attn.py >> $L 2>&1; echo "exit=$?" >> $L
tail -30 $L
