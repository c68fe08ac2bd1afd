python -m pytest tests/test_gpu_train_side.py -m gpu -x -q -k moment_records 2>&1 | tail -5
