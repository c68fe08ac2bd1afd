set -x
CMD="python bench.py --steps 30 --warmup 3 --no-cpu-baseline --e2e-steps 3 --no-rollout"
$CMD > gpurun_out/plain_r01c.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_r01c.csv $CMD > gpurun_out/ncu1_r01c.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_step --launch-skip 10 --launch-count 2 -f -o gpurun_out/prof_step_r01c $CMD > gpurun_out/ncu2_r01c.log 2>&1
tail -2 gpurun_out/ncu2_r01c.log
