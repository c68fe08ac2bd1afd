python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python tools/e2e_breakdown.py 2>&1 | grep "step()"
python bench.py --steps 500 --no-cpu-baseline --no-rollout 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e'])"
