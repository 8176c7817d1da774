mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 ) > gpurun_out/r2g_all_tests.log 2>&1
timeout 120 python __graft_entry__.py smoke > gpurun_out/r2g_smoke.log 2>&1
( time timeout 900 python bench.py > gpurun_out/r2g_bench_n1.json ) 2> gpurun_out/r2g_bench_n1.err
cat gpurun_out/r2g_all_tests.log; tail -3 gpurun_out/r2g_smoke.log; tail -3 gpurun_out/r2g_bench_n1.err
