set -x
mkdir -p gpurun_out/r02i
nvidia-smi -L
timeout 1200 python -m pytest tests -m gpu -x -q -k "tiled or c3_shape or big_table or trajectory" > gpurun_out/r02i/pytest_tiled.log 2>&1; tail -3 gpurun_out/r02i/pytest_tiled.log
timeout 1500 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r02i/pytest_multi.log 2>&1; tail -5 gpurun_out/r02i/pytest_multi.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02i/bench_n2.json 2> gpurun_out/r02i/bench_n2.err; echo "bench rc=$?"; tail -c 6000 gpurun_out/r02i/bench_n2.json; tail -5 gpurun_out/r02i/bench_n2.err
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02i/bench_n1.json 2> gpurun_out/r02i/bench_n1.err; tail -c 1500 gpurun_out/r02i/bench_n1.json
