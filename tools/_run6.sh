mkdir -p gpurun_out
L=gpurun_out/r2_c6.log
echo "=== attention forward timing experiments (dev library)" > $L
timeout 600 python tools/attn_experiments.py >> $L 2>&1; echo "exit=$?" >> $L
echo "=== sampling bench (key-parallel attn_decode)" >> $L
timeout 300 python tools/bench_sampling.py >> $L 2>&1; echo "exit=$?" >> $L
timeout 300 python -m pytest tests/test_sampling_gpu.py -q -m gpu -x >> $L 2>&1; echo "exit=$?" >> $L
tail -70 $L
