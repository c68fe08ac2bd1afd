python bench.py > gpurun_out/bench_r01_e.json 2> gpurun_out/bench_err.log; cat gpurun_out/bench_r01_e.json | cut -c1-300; tail -2 gpurun_out/bench_err.log
