#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
exec > gpurun_out/s7.log 2>&1
date
for v in base nohint diag2 diag5 diag5nh; do
  export PTG_B200_SO=$PWD/variants/$v.so
  timeout 300 python tools/microbench.py --steps 400 --no-rollout 2>&1 | tail -2
done
date
