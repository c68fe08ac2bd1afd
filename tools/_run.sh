python -m pytest tests/test_gpu_edge_cases.py -m gpu -x -q -k subproc 2>&1 | tail -12
