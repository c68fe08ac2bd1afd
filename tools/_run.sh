python -m pytest tests/test_gpu_edge_cases.py -m gpu -x -q -k "shard or full_size" 2>&1 | tail -4
python bench.py --workload ppo --steps 4 --ppo-tf32 2>/dev/null | cut -c1-900
