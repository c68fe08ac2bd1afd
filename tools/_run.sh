python tools/ppo_learning_curve.py > gpurun_out/ppo_learning_r01.json 2> gpurun_out/ppo_learning_err.log; tail -3 gpurun_out/ppo_learning_err.log; python -c "
import json; d=json.load(open(\"gpurun_out/ppo_learning_r01.json\")); print(d[\"wall_s\"]);
for r in d[\"curve\"]:
    if r[\"iteration\"]%5==0: print(r)"
