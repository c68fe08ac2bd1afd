#!/bin/bash
# round-2 GPU session 2: ncu full capture of the step kernel (default build), bench short-run check
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
exec > gpurun_out/s2.log 2>&1
set -x
date
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-strong > gpurun_out/bench_r02_b.json 2> gpurun_out/bench_r02_b.err
python -c "
import json; d=json.load(open('gpurun_out/bench_r02_b.json')); c=d['config']
print('K20', d['ms_per_step'], 'long', c['uniform_long_run']['ms_per_step'], 'sticky', c['sticky_policy']['ms_per_step'], 'roll', c['rollout_kernel']['ms_per_step'], 'copy', c['copy_gbs_this_box'], 'e2e', d['e2e']['value'])"
date
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_step --launch-skip 700 -c 2 -o gpurun_out/step_r02_full python tools/microbench.py --steps 100 --policy uniform --no-rollout > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
date
ls -la gpurun_out/*.ncu-rep
