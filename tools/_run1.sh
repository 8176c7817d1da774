mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_scan_fast_gpu.py -x -q -m gpu 2>&1 | tail -3 > gpurun_out/s20_tests.log
cat gpurun_out/s20_tests.log
timeout 300 python bench.py --steps 20 --warmup 3 --no-model --no-cpu-baseline > gpurun_out/s20_bench.json 2> gpurun_out/s20_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/s20_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['fwd_ms'], d['bwd_ms'])
for v in d['other_workloads']:
    if v['workload'] in ('vm_d192_b1','vm_d192_b2'): print(v['workload'], v['fwd_ms'], v['fwd_bwd_ms'])
PY
