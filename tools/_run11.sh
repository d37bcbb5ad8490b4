# single-GPU: two-tile forward kernel A/B
mkdir -p gpurun_out
L=gpurun_out/r2_c11.log
echo "=== attention diag with the two-tile forward (DB200_ATTN_FWD2=1)" > $L
OK2=0
if DB200_ATTN_FWD2=1 timeout 300 python tools/gpu_diag.py attn > gpurun_out/attn_fwd2.log 2>&1 && grep -q " 0 bad" gpurun_out/attn_fwd2.log; then
  OK2=1; echo "two-tile forward OK" >> $L; grep "PERF\|SUMMARY" gpurun_out/attn_fwd2.log >> $L
else
  echo "two-tile forward FAILED" >> $L; grep "BAD\|SUMMARY\|rror\|timeout" gpurun_out/attn_fwd2.log | head -20 >> $L; tail -3 gpurun_out/attn_fwd2.log >> $L
fi
echo "=== reference: current forward" >> $L
timeout 300 python tools/gpu_diag.py attn 2>&1 | grep "PERF\|SUMMARY" >> $L
if [ $OK2 = 1 ]; then
echo "=== pytest attention with the two-tile forward" >> $L
DB200_ATTN_FWD2=1 timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k attention >> $L 2>&1; echo "exit=$?" >> $L
echo "=== bench with the two-tile forward" >> $L
DB200_ATTN_FWD2=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r02_c11.json 2> gpurun_out/bench_r02_c11.err; echo "exit=$?" >> $L
python - >> $L 2>&1 <<'PY'
import json
d=None
for l in open('gpurun_out/bench_r02_c11.json'):
    if l.startswith('{'): d=json.loads(l)
if d:
    print('ms/step', d['ms_per_step'], 'tokens/s', d['value'], 'e2e', d['e2e']['value'])
    print('gemm', d['roofline']['achieved'], d['roofline']['share_of_step'])
    print('attn', {k:(round(v['tflops']),round(v['us_per_launch'])) for k,v in d['roofline_attention'].items()})
    for k,v in d.get('extra',{}).items():
        if isinstance(v, dict): print(k, v['value'], v['ms_per_step'], {kk:(round(vv['tflops'])) for kk,vv in v.get('roofline_attention',{}).items()})
PY
echo "=== ncu full on attention kernels (two-tile forward)" >> $L
DB200_ATTN_FWD2=1 timeout 500 ncu --set full --clock-control none --import-source on -k regex:attn_fwd -c 2 -o gpurun_out/prof_attn_r02h python tools/prof_attn.py >> $L 2>&1; echo "exit=$?" >> $L
fi
tail -60 $L
