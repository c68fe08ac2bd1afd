"""Does the device-side PPO loop learn?  Trains on the synthetic BS2/OP2 training split and evaluates the
deterministic policy on the validation split (the role of EvalCallback, src/rl_utils.py:456-469) every few iterations.
Writes one JSON document (iterations, env reward per step, validation return) to stdout."""
import argparse, json, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rl_ptg_b200 as ptg
from rl_ptg_b200.ppo import PPO, evaluate_policy, reference_hyper_kwargs
from rl_ptg_b200.vec_env import PtGVecEnv

ap = argparse.ArgumentParser()
ap.add_argument("--envs", type=int, default=16384)
ap.add_argument("--n-steps", type=int, default=64)
ap.add_argument("--batch", type=int, default=16384)
ap.add_argument("--epochs", type=int, default=4)
ap.add_argument("--lr", type=float, default=3e-4)
ap.add_argument("--iters", type=int, default=60)
ap.add_argument("--eval-every", type=int, default=10)
ap.add_argument("--eval-steps", type=int, default=2000)
args = ap.parse_args()
E = ptg.EnvConfiguration(scenario=2, operation="OP2")
price, op = ptg.synthetic_data(E, seed=0)
pp = ptg.Preprocessing(price, op, ptg.AgentConfiguration(), E, ptg.TrainConfiguration())
env = PtGVecEnv(pp.dict_env_kwargs("train"), args.envs, seed=3654, obs_layout="flat")
val = PtGVecEnv(pp.dict_env_kwargs("val"), 64, seed=605, obs_layout="flat")
hyper = reference_hyper_kwargs()
hyper.update(n_steps=args.n_steps, batch_size=args.batch, n_epochs=args.epochs, learning_rate=args.lr, seed=3654)
model = PPO(env, **hyper)
curve = []
t0 = time.perf_counter()
ev = evaluate_policy(model, val, args.eval_steps)
curve.append({"iteration": 0, "timesteps": 0, "val_return": ev["mean_cum_reward"]})
for it in range(1, args.iters + 1):
    model.learn(model.num_timesteps + args.envs * args.n_steps)
    row = {"iteration": it, "timesteps": model.num_timesteps, "env_reward_per_step": model.logs[-1]["env_reward_per_step"],
           "entropy": model.logs[-1]["entropy"], "value_loss": model.logs[-1]["value_loss"]}
    if it % args.eval_every == 0:
        row["val_return"] = evaluate_policy(model, val, args.eval_steps)["mean_cum_reward"]
    curve.append(row)
print(json.dumps({"config": vars(args), "wall_s": time.perf_counter() - t0, "curve": curve}))
