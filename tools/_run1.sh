mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ffn_gpu.py -x -q -m gpu 2>&1 | tail -15 > gpurun_out/s4_tests.log
timeout 900 python -m pytest tests/test_reference_gpu.py -x -q -m gpu 2>&1 | tail -8 >> gpurun_out/s4_tests.log
timeout 600 python tools/model_bench.py train fused 24 graphs 2>&1 | tail -1 > gpurun_out/s4_model.log
cat gpurun_out/s4_tests.log gpurun_out/s4_model.log
