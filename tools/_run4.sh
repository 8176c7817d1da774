mkdir -p gpurun_out
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/r2g_bench_n2.json ) 2> gpurun_out/r2g_bench_n2.err
tail -c 1500 gpurun_out/r2g_bench_n2.json; tail -4 gpurun_out/r2g_bench_n2.err
