mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_scan_fast_gpu.py tests/test_scan_gpu.py tests/test_fuzz_gpu.py -x -q -m gpu 2>&1 | tail -15 > gpurun_out/s3_tests.log
for w in vm_d192 vm_d384; do
  timeout 120 python tools/quick_bench.py $w 20 --graph 2>&1 | tail -1 | sed "s/^/main(p4) /" >> gpurun_out/s3_bench.log
  for v in fr_p0 fr_p2 fr_p6 fr_p8; do
    timeout 120 python tools/quick_bench.py $w 20 --graph --lib=variants/$v.so 2>&1 | tail -1 | sed "s/^/$v /" >> gpurun_out/s3_bench.log
  done
done
cat gpurun_out/s3_tests.log gpurun_out/s3_bench.log
