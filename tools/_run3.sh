mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 ) > gpurun_out/r2j_all_tests.log 2>&1
( time timeout 600 python bench.py > gpurun_out/r2j_bench_n1.json ) 2> gpurun_out/r2j_bench_n1.err
cat gpurun_out/r2j_all_tests.log; tail -3 gpurun_out/r2j_bench_n1.err
