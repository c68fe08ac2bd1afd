python tools/microbench.py --steps 1000 --no-rollout 2>&1 | grep -v "^$"
PTG_B200_SO=$PWD/variants/nola.so python tools/microbench.py --steps 1000 --no-rollout 2>&1 | grep -v "^$"
