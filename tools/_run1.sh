mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_reference_gpu.py tests/test_modules_gpu.py -x -q -m gpu 2>&1 | tail -3 > gpurun_out/s23_tests.log
cat gpurun_out/s23_tests.log
timeout 300 python tools/model_bench.py train fused 24 graphs 2>&1 | tail -1 | cut -c1-120
timeout 300 python tools/model_bench.py infer fused 64 graphs 2>&1 | tail -1 | cut -c1-120
