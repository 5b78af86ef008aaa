set -x
mkdir -p gpurun_out/r02q
nvidia-smi -L | wc -l
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29721 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r02q/bench_n8.json 2> gpurun_out/r02q/bench_n8.err; echo "bench rc=$?"; tail -c 4500 gpurun_out/r02q/bench_n8.json; tail -3 gpurun_out/r02q/bench_n8.err
