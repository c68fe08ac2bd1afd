#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
exec > gpurun_out/a.log 2>&1
date
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r02_1gpu.json 2> gpurun_out/bench.err; tail -2 gpurun_out/bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r02_1gpu.json')); c=d['config']
print("value %.4g ms %.5f strict %.5f e2e %.4g frac %.3f" % (d['value'], d['ms_per_step'], c['ms_per_step_without_lead_steps'], d['e2e']['value'], d['roofline']['frac']))
print("sticky", c['sticky_policy']['roofline_frac'], "biased", c['load_biased_mix']['roofline_frac'], "roll", c['rollout_kernel']['roofline_frac'], "probe", c['box_probe'])
print("flat", c['flat_layout'])
print({k:(round(v['ms_per_step']*1e3,2), round(v['roofline_frac'],3)) for k,v in c['strong_scaling']['modes'].items()})
PY
timeout 300 python tools/graphbench.py --envs 131072 --steps 2000 2>&1 | tail -2
date
