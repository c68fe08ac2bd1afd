#!/bin/bash
# round-2 GPU session 1: tests, smoke, kernel variants, bench lines of both arms
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
exec > gpurun_out/s1.log 2>&1
set -x
nvidia-smi -L
nproc
date
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
date
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
for v in default r1state nopf; do
  if [ $v = default ]; then unset PTG_B200_SO; else export PTG_B200_SO=$PWD/variants/$v.so; fi
  timeout 300 python tools/microbench.py --steps 400 2>&1 | tail -4
  timeout 300 python tools/microbench.py --steps 400 --noise off --no-rollout --policy uniform 2>&1 | tail -1
done
unset PTG_B200_SO
date
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r02_a.json 2> gpurun_out/bench_r02_a.err
tail -c 600 gpurun_out/bench_r02_a.err
date
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_r02_ref_a.json 2> gpurun_out/bench_r02_ref_a.err
date
cat gpurun_out/bench_r02_a.json gpurun_out/bench_r02_ref_a.json
