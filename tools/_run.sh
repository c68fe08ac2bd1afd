python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python __graft_entry__.py smoke 2>&1 | tail -1
python tools/e2e_breakdown.py 2>&1 | grep "step()"
