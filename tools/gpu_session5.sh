#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
exec > gpurun_out/s5.log 2>&1
date
python tools/membench.py
for v in default diag2 diag3 diag4; do
  if [ $v = default ]; then unset PTG_B200_SO; else export PTG_B200_SO=$PWD/variants/$v.so; fi
  timeout 300 python tools/microbench.py --steps 400 --no-rollout --policy sticky 2>&1 | tail -1
done
unset PTG_B200_SO
date
