#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
exec > gpurun_out/s4.log 2>&1
date
for v in default nopipe early np5 np6 diag1 diag2 default; do
  if [ $v = default ]; then unset PTG_B200_SO; else export PTG_B200_SO=$PWD/variants/$v.so; fi
  timeout 300 python tools/microbench.py --steps 400 --no-rollout 2>&1 | tail -2
done
unset PTG_B200_SO
date
