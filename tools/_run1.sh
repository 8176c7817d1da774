mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_scan_fast_gpu.py -x -q -m gpu 2>&1 | tail -4 > gpurun_out/s17_tests.log
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/s17_b1.csv python tools/prof_scan.py vm_d192_b1 3 > gpurun_out/s17_b1.log 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/s17_b2.csv python tools/prof_scan.py vm_d192_b2 3 > gpurun_out/s17_b2.log 2>&1
cat gpurun_out/s17_tests.log
python - <<'PY'
import csv
for f in ('gpurun_out/s17_b1.csv','gpurun_out/s17_b2.csv'):
    rows=[r for r in csv.reader(open(f)) if len(r)>10 and r[0].isdigit()]
    print(f)
    for r in rows[-6:]: print(r[4][:70], r[8], r[-1])
PY
timeout 300 python bench.py --steps 20 --warmup 3 --no-model --no-cpu-baseline > gpurun_out/s17_bench.json 2> gpurun_out/s17_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/s17_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['fwd_ms'], d['bwd_ms'])
for v in d['other_workloads']:
    if v['workload'] in ('vm_d192_b1','vm_d192_b2'): print(v)
PY
