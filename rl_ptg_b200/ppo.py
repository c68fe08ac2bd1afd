"""PPO on the batched GPU environment -- BASELINE.json config 5 ("end-to-end SB3 PPO training on BS2/OP2 with the
GPU VecEnv feeding the torch policy").

Stable-Baselines3 is not installed in the build container, so the reference's training call
(``PPO("MultiInputPolicy", env, learning_rate, gamma, batch_size, n_steps, gae_lambda, n_epochs,
normalize_advantage, ent_coef, policy_kwargs=dict(activation_fn, net_arch))``, ``src/rl_config_agent.py:126-149``)
is mirrored here with the same hyper-parameter names and SB3's algorithm (on-policy collection, GAE, clipped
surrogate, value loss 0.5, max_grad_norm 0.5, Adam eps 1e-5, orthogonal init), restated from SB3 2.0.0a13.

What is different is WHERE the roll-out lives: observations, actions, rewards, dones, values and log-probs never
leave the device.  Per step: ``ptg_step`` (one launch for all envs) -> ``ptg_vecnorm_*`` (VecNormalize reward,
``src/rl_utils.py:453``) -> ``ptg_features`` ([n_envs, F] rows for the policy) -> policy forward (torch).  After
``n_steps``: ``ptg_gae`` and the mini-batch updates.  The policy network is torch (library code -- it is the
reference's consumer of the path, not the path).
"""
from __future__ import annotations

import time

import numpy as np
import torch
from torch import nn

from .vec_env import PtGVecEnv
from .vec_normalize import VecNormalizeReward, feature_dim, features_tensor, gae

# config/config_agent.yaml:44-58 (PPO block)
REFERENCE_PPO_HYPER = dict(alpha=0.00005, gamma=0.973, ent_coeff=0.00001, n_steps_f=21, batch_size=203,
                           hidden_layers=2, hidden_units=358, activation="ReLU", gae_lambda=0.8002, n_epoch=13,
                           normalize_advantage=False)


class MultiInputActorCritic(nn.Module):
    """SB3 ``MultiInputActorCriticPolicy`` on the flat feature rows: CombinedExtractor (done by ``ptg_features``) ->
    separate pi / vf MLPs (``net_arch=[units]*layers``) -> action_net / value_net.  Orthogonal init like SB3
    (gain sqrt(2) for the MLPs, 0.01 for the action head, 1 for the value head)."""

    def __init__(self, n_features: int, n_actions: int = 5, hidden_layers: int = 2, hidden_units: int = 358,
                 activation: str = "ReLU", continuous: bool = False):
        super().__init__()
        act = {"ReLU": nn.ReLU, "Tanh": nn.Tanh}[activation]

        def mlp():
            layers, d = [], n_features
            for _ in range(hidden_layers):
                layers += [nn.Linear(d, hidden_units), act()]
                d = hidden_units
            return nn.Sequential(*layers)

        self.pi, self.vf = mlp(), mlp()
        self.continuous = continuous
        self.action_net = nn.Linear(hidden_units, 1 if continuous else n_actions)
        self.value_net = nn.Linear(hidden_units, 1)
        if continuous:
            self.log_std = nn.Parameter(torch.zeros(1))
        for m in list(self.pi) + list(self.vf):
            if isinstance(m, nn.Linear):
                nn.init.orthogonal_(m.weight, gain=2 ** 0.5)
                nn.init.zeros_(m.bias)
        nn.init.orthogonal_(self.action_net.weight, gain=0.01)
        nn.init.zeros_(self.action_net.bias)
        nn.init.orthogonal_(self.value_net.weight, gain=1.0)
        nn.init.zeros_(self.value_net.bias)

    def distribution(self, feat):
        out = self.action_net(self.pi(feat))
        if self.continuous:
            return torch.distributions.Normal(out.squeeze(-1), self.log_std.exp().expand(out.shape[0]))
        return torch.distributions.Categorical(logits=out)

    def value(self, feat):
        return self.value_net(self.vf(feat)).squeeze(-1)

    def forward(self, feat, deterministic: bool = False):
        dist = self.distribution(feat)
        if deterministic:
            actions = dist.mean if self.continuous else dist.probs.argmax(dim=-1)
        else:
            actions = dist.sample()
        return actions, self.value(feat), dist.log_prob(actions)

    def evaluate_actions(self, feat, actions):
        dist = self.distribution(feat)
        return self.value(feat), dist.log_prob(actions), dist.entropy()


class PPOCore:
    """Policy, optimiser, device roll-out buffers and SB3's ``PPO.train()``; how the buffers get filled is the
    subclass's business (``PPO`` below: everything on the device)."""

    def __init__(self, n_envs: int, n_features: int, device, continuous: bool = False, learning_rate: float = 5e-5,
                 n_steps: int = 4263, batch_size: int = 203, n_epochs: int = 13, gamma: float = 0.973,
                 gae_lambda: float = 0.8002, clip_range: float = 0.2, normalize_advantage: bool = False,
                 ent_coef: float = 1e-5, vf_coef: float = 0.5, max_grad_norm: float = 0.5, hidden_layers: int = 2,
                 hidden_units: int = 358, activation: str = "ReLU", seed: int | None = None):
        self.device = torch.device(device)
        self.n_envs, self.n_steps, self.batch_size, self.n_epochs = n_envs, n_steps, batch_size, n_epochs
        self.gamma, self.gae_lambda, self.clip_range = gamma, gae_lambda, clip_range
        self.normalize_advantage, self.ent_coef, self.vf_coef = normalize_advantage, ent_coef, vf_coef
        self.max_grad_norm = max_grad_norm
        if seed is not None:
            torch.manual_seed(seed)              # same initial policy on every rank
        self.continuous = continuous
        self.F = n_features
        self.policy = MultiInputActorCritic(self.F, 5, hidden_layers, hidden_units, activation,
                                            self.continuous).to(self.device)
        self.optimizer = torch.optim.Adam(self.policy.parameters(), lr=learning_rate, eps=1e-5)
        T, n, dev = n_steps, self.n_envs, self.device
        self.buf_feat = torch.empty((T, n, self.F), dtype=torch.float32, device=dev)
        self.buf_actions = torch.empty((T, n), dtype=torch.float32 if self.continuous else torch.int64, device=dev)
        self.buf_rewards = torch.empty((T, n), dtype=torch.float32, device=dev)
        self.buf_values = torch.empty((T, n), dtype=torch.float32, device=dev)
        self.buf_logp = torch.empty((T, n), dtype=torch.float32, device=dev)
        self.buf_starts = torch.empty((T, n), dtype=torch.uint8, device=dev)
        self.buf_adv = torch.empty((T, n), dtype=torch.float32, device=dev)
        self.buf_ret = torch.empty((T, n), dtype=torch.float32, device=dev)
        self.num_timesteps = 0
        self.logs: list[dict] = []
        # data-parallel training (one process per GPU, torch.distributed initialised): every rank collects from its
        # own env shard and the gradients of each mini-batch are averaged over NCCL before the optimiser step
        import torch.distributed as dist
        self._world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        if self._world > 1:
            # every rank must run the same number of mini-batches (one gradient all-reduce each) and roll-outs:
            # unequal shards (shard_range gives sizes that differ by one when n_global % world != 0) would pair
            # gradients of different epochs or hang the collective
            sizes = [None] * self._world
            dist.all_gather_object(sizes, (int(n_envs), int(n_steps), int(batch_size), int(n_epochs)))
            if len(set(sizes)) != 1:
                raise ValueError(f"data-parallel PPO needs identical (n_envs, n_steps, batch_size, n_epochs) on every "
                                 f"rank, got {sizes}: choose an env count divisible by the world size")
            for p_ in self.policy.parameters():
                dist.broadcast(p_.data, src=0)
            if seed is not None:
                torch.manual_seed(seed + 7919 * (dist.get_rank() + 1))      # independent action sampling per rank

    def _average_gradients(self):
        if self._world == 1:
            return
        import torch.distributed as dist
        grads = [p_.grad for p_ in self.policy.parameters() if p_.grad is not None]
        flat = torch.cat([g.reshape(-1) for g in grads])
        dist.all_reduce(flat)
        flat /= self._world
        off = 0
        for g in grads:
            g.copy_(flat[off:off + g.numel()].view_as(g))
            off += g.numel()

    def train(self):
        T, n = self.n_steps, self.n_envs
        total = T * n
        feat = self.buf_feat.view(total, self.F)
        actions, old_values = self.buf_actions.view(total), self.buf_values.view(total)
        old_logp, adv_all, ret_all = self.buf_logp.view(total), self.buf_adv.view(total), self.buf_ret.view(total)
        pg_losses, v_losses, ent_losses, clip_fracs = [], [], [], []
        for _ in range(self.n_epochs):
            perm = torch.randperm(total, device=self.device)
            for start in range(0, total, self.batch_size):
                idx = perm[start:start + self.batch_size]
                adv = adv_all[idx]
                if self.normalize_advantage and adv.numel() > 1:
                    adv = (adv - adv.mean()) / (adv.std() + 1e-8)
                values, logp, entropy = self.policy.evaluate_actions(feat[idx], actions[idx])
                ratio = torch.exp(logp - old_logp[idx])
                pg_loss = -torch.min(adv * ratio, adv * torch.clamp(ratio, 1 - self.clip_range, 1 + self.clip_range)).mean()
                v_loss = torch.nn.functional.mse_loss(ret_all[idx], values)
                ent_loss = -entropy.mean()
                loss = pg_loss + self.ent_coef * ent_loss + self.vf_coef * v_loss
                self.optimizer.zero_grad(set_to_none=True)
                loss.backward()
                self._average_gradients()
                nn.utils.clip_grad_norm_(self.policy.parameters(), self.max_grad_norm)
                self.optimizer.step()
                pg_losses.append(pg_loss.detach()); v_losses.append(v_loss.detach()); ent_losses.append(ent_loss.detach())
                clip_fracs.append(((ratio.detach() - 1).abs() > self.clip_range).float().mean())
        return {"pg_loss": float(torch.stack(pg_losses).mean()), "value_loss": float(torch.stack(v_losses).mean()),
                "entropy": -float(torch.stack(ent_losses).mean()), "clip_fraction": float(torch.stack(clip_fracs).mean())}


class PPO(PPOCore):
    """Same constructor vocabulary as SB3's PPO; ``env`` is a ``PtGVecEnv`` (wrapped in ``VecNormalizeReward``
    here unless ``normalize_reward=False``)."""

    def __init__(self, env: PtGVecEnv, normalize_reward: bool = True, **hyper):
        super().__init__(env.num_envs, feature_dim(env), env.device, env.action_type == "continuous", **hyper)
        self.env = env
        self.vn = VecNormalizeReward(env, gamma=0.99) if normalize_reward else None   # SB3 VecNormalize default gamma
        self._last_feat = None
        self._last_starts = torch.ones(self.n_envs, dtype=torch.uint8, device=self.device)
        self._raw_reward_sum = torch.zeros((), dtype=torch.float64, device=self.device)

    # ------------------------------------------------------------------------------------------------------
    def _reset(self):
        (self.vn or self.env).reset_tensor()
        self._last_feat = features_tensor(self.env)
        self._last_starts.fill_(1)

    @torch.no_grad()
    def collect_rollouts(self):
        if self._last_feat is None:
            self._reset()
        stepper = self.vn or self.env
        for t in range(self.n_steps):
            feat = self._last_feat
            actions, values, logp = self.policy(feat)
            self.buf_feat[t].copy_(feat)
            self.buf_actions[t].copy_(actions)
            self.buf_values[t].copy_(values)
            self.buf_logp[t].copy_(logp)
            self.buf_starts[t].copy_(self._last_starts)
            env_actions = actions.clamp(-1.0, 1.0).to(torch.float32) if self.continuous else actions
            _, reward, done = stepper.step_tensor(env_actions.contiguous())
            self._raw_reward_sum += self.env._reward.sum(dtype=torch.float64)     # unnormalised env reward [ct]
            self.buf_rewards[t].copy_(reward)
            self._last_starts.copy_(done)
            self._last_feat = features_tensor(self.env, out=self._last_feat)     # (flat layout: a view, no kernel)
        last_values = self.policy.value(self._last_feat)
        gae(self.buf_rewards, self.buf_values, self.buf_starts, last_values.contiguous(), self._last_starts,
            self.gamma, self.gae_lambda, self.buf_adv, self.buf_ret)
        self.num_timesteps += self.n_steps * self.n_envs

    def learn(self, total_timesteps: int, log_interval: int = 1, callback=None):
        """``callback``: one callable or a list (e.g. ``EvalCallback``), called as ``cb(model)`` after every iteration."""
        callbacks = [] if callback is None else (list(callback) if isinstance(callback, (list, tuple)) else [callback])
        t0 = time.perf_counter()
        it = 0
        while self.num_timesteps < total_timesteps:
            self.collect_rollouts()
            stats = self.train()
            it += 1
            torch.cuda.synchronize(self.device)
            ep = self.env.episode_stats(clear=True, reduce=False)
            stats.update(iteration=it, timesteps=self.num_timesteps, fps=self.num_timesteps / (time.perf_counter() - t0),
                         ep_rew_mean=ep["return_mean"], ep_len_mean=ep["length_mean"], episodes=ep["episodes"],
                         mean_step_reward=float(self.buf_rewards.mean()),
                         env_reward_per_step=float(self._raw_reward_sum) / (self.n_steps * self.n_envs))
            self._raw_reward_sum.zero_()
            self.logs.append(stats)
            for cb in callbacks:
                cb(self)
        return self

    @torch.no_grad()
    def predict(self, feat: torch.Tensor, deterministic: bool = True):
        return self.policy(feat, deterministic=deterministic)[0]


def evaluate_policy(model: PPO, env: PtGVecEnv, n_steps: int, deterministic: bool = True) -> dict:
    """Deterministic roll-out on a (validation / test) env, the role of ``EvalCallback`` / ``test_performance``
    (``src/rl_utils.py:456-469, 528-565``): returns the mean cumulative reward per env over ``n_steps``."""
    env.reset_tensor()
    total = torch.zeros(env.num_envs, dtype=torch.float64, device=env.device)
    feat = features_tensor(env)
    for _ in range(n_steps):
        a = model.predict(feat, deterministic)
        a = a.clamp(-1.0, 1.0).to(torch.float32) if model.continuous else a
        _, rew, _ = env.step_tensor(a.contiguous())
        total += rew.double()
        feat = features_tensor(env, out=feat)
    return {"mean_cum_reward": float(total.mean()), "std_cum_reward": float(total.std()) if env.num_envs > 1 else 0.0}


def reference_hyper_kwargs(h: dict | None = None) -> dict:
    """config_agent.yaml PPO block -> constructor kwargs (``src/rl_config_agent.py:126-149``)."""
    h = dict(REFERENCE_PPO_HYPER, **(h or {}))
    return dict(learning_rate=h["alpha"], n_steps=int(h["n_steps_f"] * h["batch_size"]), batch_size=int(h["batch_size"]),
                n_epochs=int(h["n_epoch"]), gamma=h["gamma"], gae_lambda=h["gae_lambda"],
                normalize_advantage=h["normalize_advantage"], ent_coef=h["ent_coeff"],
                hidden_layers=h["hidden_layers"], hidden_units=h["hidden_units"], activation=h["activation"])


class EvalCallback:
    """The role of SB3's ``EvalCallback`` as the reference sets it up (``src/rl_utils.py:456-469``): every
    ``eval_freq`` timesteps, roll the deterministic policy out on an evaluation env (the ``n_envs`` of the reference's
    ``eval_trials`` become the env's batch), log the mean return and keep the best policy's weights
    (``best_model_save_path`` -> ``best_state_dict`` / an optional file)."""

    def __init__(self, eval_env: PtGVecEnv, n_eval_steps: int, eval_freq: int, best_model_save_path: str | None = None,
                 deterministic: bool = True):
        self.eval_env, self.n_eval_steps, self.eval_freq = eval_env, int(n_eval_steps), int(eval_freq)
        self.best_model_save_path, self.deterministic = best_model_save_path, deterministic
        self.best_mean_reward, self.best_state_dict = -np.inf, None
        self.evaluations: list[dict] = []
        self._next = self.eval_freq

    def __call__(self, model: "PPO") -> None:
        if model.num_timesteps < self._next:
            return
        self._next = model.num_timesteps + self.eval_freq
        ev = evaluate_policy(model, self.eval_env, self.n_eval_steps, self.deterministic)
        self.evaluations.append(dict(timesteps=model.num_timesteps, **ev))
        if ev["mean_cum_reward"] > self.best_mean_reward:
            self.best_mean_reward = ev["mean_cum_reward"]
            self.best_state_dict = {k: v.detach().clone() for k, v in model.policy.state_dict().items()}
            if self.best_model_save_path:
                torch.save(self.best_state_dict, self.best_model_save_path)
