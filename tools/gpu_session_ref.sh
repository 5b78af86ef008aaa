mkdir -p gpurun_out/r03j
( time timeout 1200 python bench.py --impl reference > gpurun_out/r03j/bench_reference.json 2> gpurun_out/r03j/bench_reference.err ) 2>&1 | tail -n 3
tail -c 1200 gpurun_out/r03j/bench_reference.json
