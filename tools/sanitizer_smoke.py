"""Small end-to-end exercise of every kernel for compute-sanitizer (memcheck / racecheck / initcheck)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rl_ptg_b200 as ptg
from rl_ptg_b200.vec_env import PtGVecEnv
from rl_ptg_b200.vec_normalize import VecNormalizeReward, features_tensor, gae

E = ptg.EnvConfiguration(scenario=3, operation="OP2")
price, op = ptg.synthetic_data(E, seed=0)
kw = dict(ptg.Preprocessing(price, op, ptg.AgentConfiguration(), E, ptg.TrainConfiguration()).dict_env_kwargs("train"))
kw["eps_sim_steps"] = 30
for layout, n in (("dict", 1000), ("flat", 777)):
    env = PtGVecEnv(kw, n, seed=1, obs_layout=layout)
    vn = VecNormalizeReward(env)
    vn.reset_tensor()
    g = torch.Generator(device=env.device); g.manual_seed(0)
    for t in range(60):                                   # two auto-resets
        a = torch.randint(0, 5, (n,), generator=g, device=env.device)
        vn.step_tensor(a)
        features_tensor(env)
    env.rollout_tensor(torch.randint(0, 5, (8, n), generator=g, device=env.device))
    env.reset_tensor(mask=(np.arange(n) % 2).astype(np.uint8))
    env.step(np.zeros(n, dtype=np.int64))
    st = env.get_state(); env.set_state(st)
    print(layout, env.episode_stats())
    env.poll_error()
    env.close()
ev = PtGVecEnv(kw, 33, train_or_eval="eval", seed=2)
ev.reset(); ev.step(np.ones(33, dtype=np.int64)); ev.close()
T, n = 5, 100
z = lambda *s: torch.rand(*s, device="cuda")
gae(z(T, n), z(T, n), torch.zeros(T, n, dtype=torch.uint8, device="cuda"), z(n), torch.zeros(n, dtype=torch.uint8, device="cuda"), 0.97, 0.8)
torch.cuda.synchronize()
print("sanitizer smoke done")
