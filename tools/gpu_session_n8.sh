mkdir -p gpurun_out/r03q
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r03q/bench_n8.json 2> gpurun_out/r03q/bench_n8.err; tail -c 2500 gpurun_out/r03q/bench_n8.json
