#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
exec > gpurun_out/s6.log 2>&1
date
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge_cases.py -m gpu -x -q 2>&1 | tail -8
for v in default ws3 nows nows_nopipe default; do
  if [ $v = default ]; then unset PTG_B200_SO; else export PTG_B200_SO=$PWD/variants/$v.so; fi
  timeout 300 python tools/microbench.py --steps 400 --no-rollout 2>&1 | tail -2
done
unset PTG_B200_SO
timeout 300 python tools/microbench.py --steps 400 --no-rollout --envs 131072 2>&1 | tail -2
date
