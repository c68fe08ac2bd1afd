python -m pytest tests/test_gpu_train_side.py -m gpu -x -q -k ppo 2>&1 | tail -5
python bench.py --workload ppo --steps 4 > gpurun_out/bench_ppo_gpu.json 2> gpurun_out/bench_ppo_err.log; cat gpurun_out/bench_ppo_gpu.json; tail -3 gpurun_out/bench_ppo_err.log
python bench.py --workload ppo --ppo-envs 6 --ppo-n-steps 4263 --ppo-batch 203 --steps 1 > gpurun_out/bench_ppo_gpu_refhyper.json 2>> gpurun_out/bench_ppo_err.log; cat gpurun_out/bench_ppo_gpu_refhyper.json
python bench.py --workload ppo --impl reference --steps 1 > gpurun_out/bench_ppo_ref.json 2>> gpurun_out/bench_ppo_err.log; cat gpurun_out/bench_ppo_ref.json; tail -3 gpurun_out/bench_ppo_err.log
