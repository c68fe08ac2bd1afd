#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
exec > gpurun_out/s12.log 2>&1
date
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29533 tools/check_allreduce_stats.py 2>&1 | grep ptg_allreduce
timeout 900 $TR --master-port 29512 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench_r02_8gpu.json 2> gpurun_out/bench_r02_8gpu.err
tail -2 gpurun_out/bench_r02_8gpu.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r02_8gpu.json')); c=d['config']
print("value %.4g ms %.5f e2e %.4g" % (d['value'], d['ms_per_step'], d['e2e']['value']))
print({k:(round(v['ms_per_step']*1e3,2), round(v['roofline_frac'],3), "%.3g"%v['env_steps_per_s_total']) for k,v in c['strong_scaling']['modes'].items()})
PY
date
