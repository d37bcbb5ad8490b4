# 8-GPU re-measure with the SM reservation (charged 8x: weak + strong only)
mkdir -p gpurun_out
L=gpurun_out/r2_mg8b.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541"
echo "=== smoke" > $L
if timeout 120 $TR tools/dp_smoke.py > gpurun_out/dp_smoke8.log 2>&1 && grep -q "DP SMOKE OK" gpurun_out/dp_smoke8.log; then echo "smoke OK" >> $L; else echo "smoke FAILED" >> $L; tail -5 gpurun_out/dp_smoke8.log >> $L; tail $L; exit 1; fi
echo "=== bench N=8 weak" >> $L
timeout 200 $TR bench.py --gpus 8 --steps 10 --warmup 3 --no-extra > gpurun_out/bench_r02_n8_weak_b.json 2> gpurun_out/bench_r02_n8_weak_b.err; echo "exit=$?" >> $L
echo "=== bench N=8 strong" >> $L
timeout 200 $TR bench.py --gpus 8 --steps 10 --warmup 3 --no-extra --scaling strong > gpurun_out/bench_r02_n8_strong_b.json 2> gpurun_out/bench_r02_n8_strong_b.err; echo "exit=$?" >> $L
echo "=== bench N=1 on the same box" >> $L
timeout 200 python bench.py --steps 10 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/bench_r02_n1_samebox8.json 2> /dev/null; echo "exit=$?" >> $L
python - >> $L 2>&1 <<'PY'
import json
for n in ('bench_r02_n8_weak_b','bench_r02_n8_strong_b','bench_r02_n1_samebox8'):
    d=None
    try:
        for l in open(f'gpurun_out/{n}.json'):
            if l.startswith('{'): d=json.loads(l)
    except OSError: pass
    if d:
        r=d['roofline']
        print(n, 'ms', round(d['ms_per_step'],2), 'tok/s', round(d['value']), 'gemm', round(r['achieved']), 'attn', {k:(round(v['tflops']),round(v['us_per_launch'])) for k,v in d['roofline_attention'].items()})
PY
tail -30 $L
