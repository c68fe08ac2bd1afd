#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
exec > gpurun_out/s3.log 2>&1
set -x
date
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
for v in default nopipe; do
  if [ $v = default ]; then unset PTG_B200_SO; else export PTG_B200_SO=$PWD/variants/$v.so; fi
  timeout 300 python tools/microbench.py --steps 400 2>&1 | tail -4
  timeout 300 python tools/microbench.py --steps 400 --layout flat --no-rollout 2>&1 | tail -2
done
unset PTG_B200_SO
for c in 3 5; do
  PTG_CTAS_PER_SM=$c timeout 300 python tools/microbench.py --steps 400 --no-rollout 2>&1 | tail -2
done
timeout 300 python tools/microbench.py --steps 400 --no-rollout --envs 131072 2>&1 | tail -2
PTG_B200_SO=$PWD/variants/nopipe.so timeout 300 python tools/microbench.py --steps 400 --no-rollout --envs 131072 2>&1 | tail -2
date
