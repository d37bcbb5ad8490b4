mkdir -p gpurun_out
timeout 170 python bench.py --no-extra > gpurun_out/bench_r02_n1_final.json 2> gpurun_out/bench_r02_n1_final.err
grep "^{" gpurun_out/bench_r02_n1_final.json | cut -c1-300
