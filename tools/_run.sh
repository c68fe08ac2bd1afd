python -m pytest tests -m gpu -x -q 2>&1 | tail -4; python tools/microbench.py --steps 1000 2>&1 | grep -v "^$"
