mkdir -p gpurun_out/r03u
timeout 1500 python -m pytest tests/test_gpu_multi.py -m gpu -x -q --durations=5 > gpurun_out/r03u/pytest_multi_n2.log 2>&1; echo "rc=$?" >> gpurun_out/r03u/pytest_multi_n2.log; tail -n 10 gpurun_out/r03u/pytest_multi_n2.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r03u/bench_n2.json 2> gpurun_out/r03u/bench_n2.err; tail -c 600 gpurun_out/r03u/bench_n2.json
