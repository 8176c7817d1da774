mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 ) > gpurun_out/r2h_all_tests.log 2>&1
timeout 120 python __graft_entry__.py smoke > gpurun_out/r2h_smoke.log 2>&1
( time timeout 900 python bench.py > gpurun_out/r2h_bench_n1.json ) 2> gpurun_out/r2h_bench_n1.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"carry|fixup|scan_fwdr" -f -o gpurun_out/r2h_seg_b1 python tools/prof_scan.py vm_d192_b1 1 > gpurun_out/r2h_ncu_seg.log 2>&1
timeout 120 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2h_bench_ref.json 2> gpurun_out/r2h_bench_ref.err
cat gpurun_out/r2h_all_tests.log; tail -3 gpurun_out/r2h_smoke.log; tail -3 gpurun_out/r2h_bench_n1.err; tail -2 gpurun_out/r2h_ncu_seg.log; cut -c1-400 gpurun_out/r2h_bench_ref.json
