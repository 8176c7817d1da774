mkdir -p gpurun_out
( time timeout 900 python bench.py > gpurun_out/r2f_bench_n1.json ) 2> gpurun_out/r2f_bench_n1.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2f_launches.csv python bench.py --steps 3 --warmup 3 --no-model --no-extras --no-cpu-baseline > gpurun_out/r2f_ncu_bench.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -f -o gpurun_out/r2f_vm_d192 python tools/prof_scan.py vm_d192 1 > gpurun_out/r2f_ncu_scan.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:dwnhwc -s 7 -c 7 -f -o gpurun_out/r2f_ffn python tools/prof_ffn.py > gpurun_out/r2f_ncu_ffn.log 2>&1
tail -c 3000 gpurun_out/r2f_bench_n1.json; tail -5 gpurun_out/r2f_bench_n1.err; ls -la gpurun_out
