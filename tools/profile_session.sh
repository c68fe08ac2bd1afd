#!/bin/bash
# Round profile artefacts (run on a GPU box; outputs land in gpurun_out/, summarised by tools/summarize_profile.py):
#   1. both bench arms WITHOUT ncu (the numbers)   2. ncu launch list of the bench command   3. ncu --set full of k_step
cd "$(dirname "$0")/.."
TAG=${1:-r02}
mkdir -p gpurun_out
exec > gpurun_out/profile_$TAG.log 2>&1
set -x
date
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_${TAG}_1gpu.json 2> gpurun_out/bench_${TAG}_1gpu.err || tail -5 gpurun_out/bench_${TAG}_1gpu.err
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_${TAG}_reference_arm.json 2> gpurun_out/bench_${TAG}_reference_arm.err
date
CMD="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-strong --no-ppo-line --no-rollout --e2e-steps 3 --preroll-steps 64 --presteps 64"
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launches_$TAG.log 2>&1
tail -2 gpurun_out/ncu_launches_$TAG.log
date
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_step --launch-skip 700 -c 2 -o gpurun_out/step_${TAG}_full python tools/microbench.py --steps 100 --policy uniform --no-rollout > gpurun_out/ncu_full_$TAG.log 2>&1
tail -2 gpurun_out/ncu_full_$TAG.log
date
python tools/membench.py
