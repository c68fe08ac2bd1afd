for v in l2h l2n; do PTG_B200_SO=$PWD/variants/$v.so python tools/microbench.py --steps 500 --no-rollout 2>&1 | grep -v "^$"; done
