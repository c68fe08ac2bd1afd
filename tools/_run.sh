python tools/microbench.py --steps 1000 2>&1 | grep -v "^$"
PTG_B200_SO=$PWD/variants/nopdl.so python tools/microbench.py --steps 1000 2>&1 | grep -v "^$"
python tools/graphbench.py --envs 1048576 2>&1 | grep -v "^$" | tail -2
python tools/graphbench.py --envs 131072 2>&1 | grep -v "^$" | tail -2
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
