mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_ffn_gpu.py -x -q -m gpu 2>&1 | tail -3 > gpurun_out/s12_ffn_tests.log
timeout 200 python tools/bench_ffn.py bf16 > gpurun_out/s12_bench_ffn.log 2>&1
cat gpurun_out/s12_ffn_tests.log gpurun_out/s12_bench_ffn.log
