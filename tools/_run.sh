python -m pytest tests/test_gpu_train_side.py -m gpu -x -q -k calculate_optimum 2>&1 | tail -8
