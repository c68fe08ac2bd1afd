python tools/trainside_bench.py 2>&1 | grep -v "^$" | tail -5 | head -3
python -m pytest tests/test_gpu_train_side.py -m gpu -x -q 2>&1 | tail -3
