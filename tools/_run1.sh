mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_ffn_gpu.py -x -q -m gpu 2>&1 | tail -5 > gpurun_out/s11_ffn_tests.log
timeout 200 python tools/bench_ffn.py bf16 > gpurun_out/s11_bench_ffn.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:dwnhwc -s 7 -c 7 -f -o gpurun_out/s11_ffn python tools/prof_ffn.py > gpurun_out/s11_ncu_ffn.log 2>&1
cat gpurun_out/s11_ffn_tests.log gpurun_out/s11_bench_ffn.log; tail -3 gpurun_out/s11_ncu_ffn.log
