mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_scan_gpu.py tests/test_fuzz_gpu.py tests/test_scan_fast_gpu.py -x -q -m gpu 2>&1 | tail -5 > gpurun_out/s3_tests.log
for v in fr_nox fr_u1 fr_u4 fr_r2; do
  timeout 120 python tools/quick_bench.py vm_d192 20 --graph --lib=variants/$v.so 2>&1 | tail -1 | sed "s/^/$v /" >> gpurun_out/s3_bench.log
done
for w in vm_d192 vm_d96 vm_d384 vm_d192_b2; do
  timeout 120 python tools/quick_bench.py $w 20 --graph 2>&1 | tail -1 | sed "s/^/main /" >> gpurun_out/s3_bench.log
done
cat gpurun_out/s3_tests.log gpurun_out/s3_bench.log
