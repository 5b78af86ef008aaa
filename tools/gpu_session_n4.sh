mkdir -p gpurun_out/r03v
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/r03v/bench_n4.json 2> gpurun_out/r03v/bench_n4.err; tail -c 400 gpurun_out/r03v/bench_n4.json
