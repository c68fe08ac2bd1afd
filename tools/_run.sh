python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python __graft_entry__.py smoke 2>&1 | tail -2
python bench.py > gpurun_out/bench_r01_c.json 2> gpurun_out/bench_err.log; cat gpurun_out/bench_r01_c.json; tail -5 gpurun_out/bench_err.log
