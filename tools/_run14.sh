mkdir -p gpurun_out
L=gpurun_out/r2_c14.log
echo "=== smoke()" > $L
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" >> $L 2>&1; echo "exit=$?" >> $L
echo "=== tokenizer parity test + kernel tests" >> $L
timeout 600 python -m pytest tests/test_parity_baseline_gpu.py -q -m gpu -k "tokenizer" >> $L 2>&1; echo "exit=$?" >> $L
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_sampling_gpu.py -q -m gpu >> $L 2>&1; echo "exit=$?" >> $L
echo "=== bench (traffic field)" >> $L
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('ms', d['ms_per_step'], 'traffic', d['roofline']['traffic'], 'launches', d['gpu_launches'])" >> $L 2>&1
tail -30 $L
