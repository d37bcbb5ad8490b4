mkdir -p gpurun_out
L=gpurun_out/r2_c15.log
echo "=== A: producer-side fence, loads after the fence" > $L
timeout 300 python -m pytest tests/test_engine_gpu.py -q -m gpu -k "conv or vae" >> $L 2>&1; echo "exit=$?" >> $L
timeout 200 python bench.py --vae-example --steps 50 --warmup 10 2>/dev/null | cut -c1-160 >> $L
echo "=== B: consumer-side fence (A/B build)" >> $L
export DB200_LIB=$PWD/dalle_mtf_b200/libdalle_b200_ab.so
timeout 300 python -m pytest tests/test_engine_gpu.py -q -m gpu -k "conv or vae" >> $L 2>&1; echo "exit=$?" >> $L
timeout 200 python bench.py --vae-example --steps 50 --warmup 10 2>/dev/null | cut -c1-160 >> $L
timeout 200 python bench.py --vae-example --steps 50 --warmup 10 2>/dev/null | cut -c1-160 >> $L
unset DB200_LIB
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_vae_example_r02d.csv python bench.py --vae-example --steps 1 --warmup 1 > /dev/null 2>&1; echo "exit=$?" >> $L
tail -30 $L
