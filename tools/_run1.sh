mkdir -p gpurun_out
for cfg in "8 4" "8 8" "12 12" "6 6" "4 4"; do
  set -- $cfg
  timeout 200 python bench.py --steps 20 --warmup 3 --no-model --no-extras --no-cpu-baseline --e2e-chunks $1 --e2e-streams $2 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('chunks/streams', '$1/$2', 'e2e', d['e2e']['value'], 'value', d['value'])"
done
