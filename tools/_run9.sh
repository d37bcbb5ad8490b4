# single-GPU validation of the persistent / pipelined attention kernels (with fallback), fp32 conv pipeline, bench
mkdir -p gpurun_out
L=gpurun_out/r2_c9.log
echo "=== attention sanity gate (persistent kernels); falls back to the previous build when it fails" > $L
GATE=new
for mode in 3 1 0; do
  if DB200_ATTN_PERSIST=$mode timeout 300 python tools/gpu_diag.py attn > gpurun_out/attn_gate_$mode.log 2>&1 && grep -q " 0 bad" gpurun_out/attn_gate_$mode.log; then
    echo "attention OK with DB200_ATTN_PERSIST=$mode" >> $L; grep "PERF\|SUMMARY" gpurun_out/attn_gate_$mode.log >> $L
    export DB200_ATTN_PERSIST=$mode; break
  else
    echo "attention FAILED with DB200_ATTN_PERSIST=$mode" >> $L; grep "BAD\|SUMMARY\|rror" gpurun_out/attn_gate_$mode.log | head -20 >> $L; tail -3 gpurun_out/attn_gate_$mode.log >> $L
    if [ $mode = 0 ]; then GATE=prev; export DB200_LIB=$PWD/dalle_mtf_b200/libdalle_b200_prev.so; unset DB200_ATTN_PERSIST; echo "using libdalle_b200_prev.so" >> $L; fi
  fi
done
echo "=== kernel + engine tests" >> $L
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_engine_gpu.py -q -m gpu >> $L 2>&1; echo "exit=$?" >> $L
echo "=== bench" >> $L
DB200_BENCH_VERBOSE=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r02_c9.json 2> gpurun_out/bench_r02_c9.err; echo "exit=$?" >> $L
grep "^{" gpurun_out/bench_r02_c9.json | cut -c1-300 >> $L
python - >> $L 2>&1 <<'PY'
import json
d=None
for l in open('gpurun_out/bench_r02_c9.json'):
    if l.startswith('{'): d=json.loads(l)
if d:
    print('ms/step', d['ms_per_step'], 'tokens/s', d['value'], 'e2e', d['e2e']['value'])
    print('gemm', d['roofline']['achieved'], d['roofline']['share_of_step'])
    print('attn', {k:(round(v['tflops']),round(v['us_per_launch'])) for k,v in d['roofline_attention'].items()})
    print('vae', d['vae']['value'], d['vae']['ms_per_step'])
    for k,v in d.get('extra',{}).items():
        if isinstance(v, dict): print(k, v['value'], v['ms_per_step'], {kk:(round(vv['tflops'])) for kk,vv in v.get('roofline_attention',{}).items()})
PY
if [ $GATE = new ]; then
echo "=== attention variants (dev library)" >> $L
timeout 500 python tools/attn_experiments.py >> $L 2>&1; echo "exit=$?" >> $L
echo "=== attention timeline (dev library)" >> $L
DB200_LIB=$PWD/dalle_mtf_b200/libdalle_b200_dev.so timeout 200 python tools/attn_trace.py > gpurun_out/attn_trace_r02_persist.txt 2>&1; echo "exit=$?" >> $L
fi
echo "=== vae_example" >> $L
timeout 200 python bench.py --vae-example --steps 50 --warmup 10 2>/dev/null | cut -c1-200 >> $L; echo "exit=$?" >> $L
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_vae_example_r02b.csv python bench.py --vae-example --steps 1 --warmup 1 > /dev/null 2>&1; echo "exit=$?" >> $L
tail -120 $L
