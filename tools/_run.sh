python tools/microbench.py --steps 1000 --layout flat 2>&1 | grep -v "^$"
python tools/microbench.py --steps 1000 --no-rollout 2>&1 | grep -v "^$"
