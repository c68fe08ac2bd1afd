#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
exec > gpurun_out/io.log 2>&1
date
python tools/membench.py | head -3
for v in default io5 io6 io7 io8 io9; do
  if [ $v = default ]; then unset PTG_B200_SO; else export PTG_B200_SO=$PWD/variants/$v.so; fi
  timeout 300 python tools/microbench.py --steps 400 --no-rollout --policy sticky 2>&1 | tail -1
done
date
