# single-GPU: persistent attention (two-buffer forward), fp32 conv with MMA warp, bench, trace, ncu attention capture
mkdir -p gpurun_out
L=gpurun_out/r2_c10.log
echo "=== attention sanity gate" > $L
GATE=new
if timeout 300 python tools/gpu_diag.py attn > gpurun_out/attn_gate.log 2>&1 && grep -q " 0 bad" gpurun_out/attn_gate.log; then
  echo "attention OK" >> $L; grep "PERF\|SUMMARY" gpurun_out/attn_gate.log >> $L
else
  echo "attention FAILED" >> $L; grep "BAD\|SUMMARY\|rror" gpurun_out/attn_gate.log | head -20 >> $L; tail -3 gpurun_out/attn_gate.log >> $L
  GATE=prev; export DB200_LIB=$PWD/dalle_mtf_b200/libdalle_b200_prev.so; echo "using libdalle_b200_prev.so" >> $L
fi
echo "=== kernel + engine tests" >> $L
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_engine_gpu.py -q -m gpu >> $L 2>&1; echo "exit=$?" >> $L
echo "=== bench" >> $L
DB200_BENCH_VERBOSE=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r02_c10.json 2> gpurun_out/bench_r02_c10.err; echo "exit=$?" >> $L
python - >> $L 2>&1 <<'PY'
import json
d=None
for l in open('gpurun_out/bench_r02_c10.json'):
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
timeout 400 python tools/attn_experiments.py >> $L 2>&1; echo "exit=$?" >> $L
echo "=== attention timeline (dev library)" >> $L
DB200_LIB=$PWD/dalle_mtf_b200/libdalle_b200_dev.so timeout 200 python tools/attn_trace.py > gpurun_out/attn_trace_r02_final.txt 2>&1; echo "exit=$?" >> $L
echo "=== ncu full on attention kernels" >> $L
timeout 500 ncu --set full --clock-control none --import-source on -k regex:attn_ -c 8 -o gpurun_out/prof_attn_r02g python tools/prof_attn.py >> $L 2>&1; echo "exit=$?" >> $L
fi
echo "=== vae_example" >> $L
timeout 200 python bench.py --vae-example --steps 50 --warmup 10 2>/dev/null | cut -c1-200 >> $L; echo "exit=$?" >> $L
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_vae_example_r02c.csv python bench.py --vae-example --steps 1 --warmup 1 > /dev/null 2>&1; echo "exit=$?" >> $L
tail -100 $L
