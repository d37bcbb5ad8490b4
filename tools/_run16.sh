mkdir -p gpurun_out
L=gpurun_out/r2_c16.log
echo "=== full pytest -m gpu (final tree)" > $L
DB200_PARITY_LOG=gpurun_out/parity_r02_recheck.jsonl timeout 1100 python -m pytest tests -q -m gpu -x > gpurun_out/pytest_gpu_r02_final.log 2>&1; echo "exit=$?" >> $L; tail -5 gpurun_out/pytest_gpu_r02_final.log >> $L
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" >> $L 2>&1
tail -12 $L
