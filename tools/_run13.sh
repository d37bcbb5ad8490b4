# 2-GPU: lazy communicator creation + per-workload CTA cap
mkdir -p gpurun_out
L=gpurun_out/r2_c13.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551"
echo "=== smoke" > $L
timeout 150 $TR tools/dp_smoke.py > gpurun_out/dp_smoke2.log 2>&1; echo "exit=$?" >> $L; grep "SMOKE" gpurun_out/dp_smoke2.log >> $L
echo "=== pytest 2-GPU" >> $L
timeout 400 python -m pytest tests/test_dp_gpu.py -q -m gpu >> $L 2>&1; echo "exit=$?" >> $L
echo "=== bench N=2 weak" >> $L
timeout 300 $TR bench.py --gpus 2 --steps 10 --warmup 3 --no-extra > gpurun_out/bench_r02_n2_weak_d.json 2> gpurun_out/bench_r02_n2_weak_d.err; echo "exit=$?" >> $L
echo "=== bench N=2 strong" >> $L
timeout 300 $TR bench.py --gpus 2 --steps 10 --warmup 3 --no-extra --scaling strong > gpurun_out/bench_r02_n2_strong_d.json 2> gpurun_out/bench_r02_n2_strong_d.err; echo "exit=$?" >> $L
python - >> $L 2>&1 <<'PY'
import json
for n in ('bench_r02_n2_weak_d','bench_r02_n2_strong_d'):
    d=None
    try:
        for l in open(f'gpurun_out/{n}.json'):
            if l.startswith('{'): d=json.loads(l)
    except OSError: pass
    if d:
        r=d['roofline']
        print(n, 'ms', round(d['ms_per_step'],2), 'tok/s', round(d['value']), 'gemm', round(r['achieved']), 'attn', {k:(round(v['tflops']),round(v['us_per_launch'])) for k,v in d['roofline_attention'].items()})
PY
tail -5 gpurun_out/bench_r02_n2_weak_d.err >> $L
tail -30 $L
