mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 ) > gpurun_out/r2i_all_tests.log 2>&1
( time timeout 900 python bench.py > gpurun_out/r2i_bench_n1.json ) 2> gpurun_out/r2i_bench_n1.err
timeout 300 python tools/prof_model.py fused capturable > gpurun_out/r2i_torchprof_final.txt 2> gpurun_out/r2i_torchprof.err
cat gpurun_out/r2i_all_tests.log; tail -3 gpurun_out/r2i_bench_n1.err; head -12 gpurun_out/r2i_torchprof_final.txt
