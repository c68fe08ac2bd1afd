"""-m gpu tests of the callers either side of the env path: VecNormalize reward
normalisation, flat MultiInputPolicy features and GAE on the device, against the numpy restatement of the
Stable-Baselines3 algorithms in oracle/sb3_restated.py (parity unpinned: SB3 itself is not installed here).

Tolerances: VecNormalize statistics 1e-10 relative (fp64, different but deterministic summation tree);
normalised rewards 1e-6 relative (fp32 output); features bit-exact; GAE 1e-6 relative (fp32 buffers)."""
import numpy as np
import pytest

from helpers import synthetic_kwargs

pytestmark = pytest.mark.gpu


def _env(n, overrides=None, **kw):
    from rl_ptg_b200.vec_env import PtGVecEnv
    return PtGVecEnv(synthetic_kwargs(overrides or dict(scenario=2, operation="OP2")), n, seed=3654, **kw)


@pytest.mark.parametrize("n_envs", [6, 1000, 70001])
def test_vecnormalize_reward_matches_sb3_restatement(n_envs):
    import torch
    from oracle.sb3_restated import VecNormalizeRewardRef
    from rl_ptg_b200.vec_normalize import VecNormalizeReward
    env = _env(n_envs)
    vn = VecNormalizeReward(env, gamma=0.99, clip_reward=10.0)
    ref = VecNormalizeRewardRef(n_envs, gamma=0.99, clip_reward=10.0)
    vn.reset_tensor()
    g = torch.Generator(device=env.device); g.manual_seed(7)
    for t in range(40):
        a = torch.randint(0, 5, (n_envs,), generator=g, device=env.device)
        _, nrew, done = vn.step_tensor(a)
        raw = env._reward.cpu().numpy()
        want = ref.step(raw.astype(np.float32), done.cpu().numpy().astype(bool))
        got = nrew.cpu().numpy()
        assert np.allclose(got, want, rtol=1e-6, atol=1e-9), f"step {t}"
        assert np.isclose(vn.ret_rms.mean, ref.ret_rms.mean, rtol=1e-10, atol=1e-13)
        assert np.isclose(vn.ret_rms.var, ref.ret_rms.var, rtol=1e-10)
        assert np.isclose(vn.ret_rms.count, ref.ret_rms.count, rtol=1e-14)
        assert np.allclose(vn.returns.cpu().numpy(), ref.returns, rtol=1e-12, atol=1e-12)
    # evaluation mode: statistics frozen, normalisation still applied (SB3: env.training = False)
    vn.training = ref.training = False
    before = (vn.ret_rms.mean, vn.ret_rms.var, vn.ret_rms.count)
    a = torch.randint(0, 5, (n_envs,), generator=g, device=env.device)
    _, nrew, done = vn.step_tensor(a)
    want = ref.step(env._reward.cpu().numpy(), done.cpu().numpy().astype(bool))
    assert np.allclose(nrew.cpu().numpy(), want, rtol=1e-6, atol=1e-9)
    assert before == (vn.ret_rms.mean, vn.ret_rms.var, vn.ret_rms.count)
    env.close()


def test_vecnormalize_is_bit_reproducible():
    import torch
    from rl_ptg_b200.vec_normalize import VecNormalizeReward
    outs = []
    for _ in range(2):
        env = _env(50000)
        vn = VecNormalizeReward(env)
        vn.reset_tensor()
        g = torch.Generator(device=env.device); g.manual_seed(3)
        for t in range(10):
            _, r, _ = vn.step_tensor(torch.randint(0, 5, (50000,), generator=g, device=env.device))
        outs.append((r.cpu().numpy().copy(), vn.ret_rms.mean, vn.ret_rms.var))
        env.close()
    assert np.array_equal(outs[0][0], outs[1][0]) and outs[0][1:] == outs[1][1:]


@pytest.mark.parametrize("overrides", [dict(scenario=2, operation="OP2"),
                                       dict(scenario=1, operation="OP1", raw_modified="raw"),
                                       dict(scenario=3, operation="OP2", price_ahead=6),
                                       dict(scenario=1, operation="OP2", raw_modified="raw", price_ahead=16)])
@pytest.mark.parametrize("n_envs", [5, 257, 4099])
def test_flat_features_match_combined_extractor(overrides, n_envs):
    import torch
    from oracle.sb3_restated import combined_extractor_ref
    from rl_ptg_b200.vec_normalize import feature_dim, feature_names, features_tensor
    env = _env(n_envs, overrides)
    obs = env.reset()
    F = feature_dim(env)
    assert len(feature_names(env)) == F
    g = torch.Generator(device=env.device); g.manual_seed(1)
    for t in range(6):
        want = combined_extractor_ref(obs)
        got = features_tensor(env).cpu().numpy()
        assert got.shape == (n_envs, F) == want.shape
        assert np.array_equal(got, want), f"features differ at step {t}"
        a = torch.randint(0, 5, (n_envs,), generator=g, device=env.device).cpu().numpy()
        obs, _, _, _ = env.step(a)
    env.close()


@pytest.mark.parametrize("T,n", [(1, 7), (5, 1000), (64, 5003)])
def test_gae_matches_sb3_restatement(T, n):
    import torch
    from oracle.sb3_restated import gae_ref
    from rl_ptg_b200.vec_normalize import gae
    rng = np.random.default_rng(T * 1000 + n)
    rewards = rng.normal(0, 1, (T, n)).astype(np.float32)
    values = rng.normal(0, 2, (T, n)).astype(np.float32)
    starts = (rng.random((T, n)) < 0.05).astype(np.float32)
    last_values = rng.normal(0, 2, n).astype(np.float32)
    dones = rng.random(n) < 0.1
    gamma, lam = 0.973, 0.8002            # config/config_agent.yaml:46,53
    want_adv, want_ret = gae_ref(rewards, values, starts, last_values, dones, gamma, lam)
    dev = "cuda:0"
    adv, ret = gae(torch.from_numpy(rewards).to(dev), torch.from_numpy(values).to(dev),
                   torch.from_numpy(starts.astype(np.uint8)).to(dev), torch.from_numpy(last_values).to(dev),
                   torch.from_numpy(dones.astype(np.uint8)).to(dev), gamma, lam)
    assert np.allclose(adv.cpu().numpy(), want_adv, rtol=1e-6, atol=1e-6)
    assert np.allclose(ret.cpu().numpy(), want_ret, rtol=1e-6, atol=1e-6)
    # evaluation order and dtypes follow numpy's: expect (near) bit equality, report the fraction that is exact
    exact = np.mean(adv.cpu().numpy() == want_adv)
    assert exact > 0.999, f"only {exact:.4f} of the advantages are bit-identical"


def test_ppo_rollout_buffers_are_consistent():
    """Two PPO iterations on the device: the roll-out buffers obey the SB3 definitions (episode_starts shifted dones,
    advantages = GAE of the stored rewards / values, stored rewards = VecNormalize of the env rewards) and the update
    moves the policy."""
    import torch
    from oracle.sb3_restated import gae_ref
    from rl_ptg_b200.ppo import PPO, reference_hyper_kwargs
    env = _env(512)
    hyper = reference_hyper_kwargs()
    hyper.update(n_steps=16, batch_size=1024, n_epochs=2, seed=1)
    model = PPO(env, **hyper)
    before = [p.detach().clone() for p in model.policy.parameters()]
    model.collect_rollouts()
    last_values = model.policy.value(model._last_feat).detach().cpu().numpy()
    adv, ret = gae_ref(model.buf_rewards.cpu().numpy(), model.buf_values.cpu().numpy(),
                       model.buf_starts.cpu().numpy().astype(np.float32), last_values,
                       model._last_starts.cpu().numpy().astype(bool), model.gamma, model.gae_lambda)
    assert np.allclose(model.buf_adv.cpu().numpy(), adv, rtol=1e-6, atol=1e-6)
    assert np.allclose(model.buf_ret.cpu().numpy(), ret, rtol=1e-6, atol=1e-6)
    assert model.buf_starts[0].all() and not model.buf_starts[1:].any()      # no episode ends within 16 steps
    assert torch.isfinite(model.buf_feat).all() and model.buf_feat.shape == (16, 512, 40)
    onehot = model.buf_feat[..., 5:11]
    assert torch.equal(onehot.sum(-1), torch.ones_like(onehot[..., 0]))
    assert (model.buf_rewards.abs() <= 10.0).all()                            # VecNormalize clip_reward
    stats = model.train()
    assert all(np.isfinite(v) for v in stats.values())
    assert any(not torch.equal(a, b) for a, b in zip(before, model.policy.parameters()))
    model.learn(model.num_timesteps + 16 * 512)
    assert model.logs and np.isfinite(model.logs[-1]["fps"])
    env.close()


def test_postprocessing_matches_reference_info_rows():
    """Postprocessing.test_performance (src/rl_utils.py:528-565) on the real BS2/OP2 test split, replaying the golden
    action sequence as the 'policy': the (steps, 24) statistics equal the reference's eval-mode info rows, and the
    batched device roll-out gives the same columns for every env."""
    import torch
    from helpers import golden_kwargs, load_golden
    from rl_ptg_b200._abi import STATE_NAMES
    from rl_ptg_b200.postprocessing import Postprocessing, eval_stats_tensor
    from rl_ptg_b200.vec_env import PtGVecEnv
    g = load_golden("bs2_op2_eval_test")
    kw = golden_kwargs("bs2_op2_eval_test")
    steps = int(kw["eps_sim_steps"])                 # 8640: five steps into the second episode, like the reference

    class Replay:
        def __init__(self):
            self.t = 0

        def predict(self, obs, deterministic=True):
            a = g["actions"][self.t]
            self.t += 1
            return a, None

    env = PtGVecEnv(kw, 1, train_or_eval="eval", seed=g["meta"]["seed"])
    pp = Postprocessing(env, Replay(), steps)
    pp.test_performance()
    got = np.stack([pp.stats_dict_test[name] for name in pp.stats_names], axis=1)
    want = g["infos"][:steps, 0, :].copy()
    term = g["ints"][:steps, 0, 4].astype(bool)
    want[term] = 0.0                                  # `if not terminated` leaves those rows zero
    assert np.allclose(got, want, rtol=1e-9, atol=1e-12)
    assert list(pp.stats_dict_test) == pp.stats_names and len(pp.stats_names) == 24
    env.close()

    # batched on the device: 64 envs with the same seed-independent action tape (noise differs per env, prices do not)
    n = 64
    envb = PtGVecEnv(kw, n, train_or_eval="eval", seed=g["meta"]["seed"])
    acts = torch.from_numpy(g["actions"][:steps, 0].astype(np.int64)).to(envb.device)
    tcount = [0]

    def policy(obs):
        a = acts[tcount[0]].expand(n)
        tcount[0] += 1
        return a

    stats = eval_stats_tensor(envb, policy, 300).cpu().numpy()
    assert stats.shape == (300, 24, n)
    assert np.allclose(stats[:, :, 0], want[:300], rtol=1e-9, atol=1e-12)      # env 0 has the golden seed
    assert np.array_equal(stats[:, 0, :], np.tile(np.arange(300.0)[:, None], (1, n)))   # "step" column
    assert np.allclose(stats[:, 1, :], stats[:, 1, :1])                        # same prices for every env
    envb.close()


def test_vecnormalize_moment_records_combine_like_one_batch():
    """reduce="global" without NCCL: two half-batches produce one 24-byte moment record each; folding both records
    (rank order) into the statistics gives what one batch over all envs gives -- so the reward statistics do not
    depend on the sharding."""
    import ctypes as C
    import torch
    from rl_ptg_b200 import _lib
    from rl_ptg_b200.vec_env import PtGVecEnv
    L = _lib.load()
    n, dev = 6000, "cuda:0"
    whole = _env(n)
    halves = [PtGVecEnv(synthetic_kwargs(dict(scenario=2, operation="OP2")), n // 2, seed=3654, env_id_offset=q * (n // 2),
                        n_envs_global=n) for q in range(2)]
    g = torch.Generator(device=dev); g.manual_seed(11)
    p = PtGVecEnv._ptr
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    st = {k: [torch.tensor([0.0, 1.0, 1e-4, 0.0], dtype=torch.float64, device=dev) for _ in range(2)] for k in "wab"}
    ret = {"w": torch.zeros(n, dtype=torch.float64, device=dev), "a": torch.zeros(n // 2, dtype=torch.float64, device=dev),
           "b": torch.zeros(n // 2, dtype=torch.float64, device=dev)}
    out = {"w": torch.zeros(n, device=dev), "a": torch.zeros(n // 2, device=dev), "b": torch.zeros(n // 2, device=dev)}
    for env in [whole] + halves:
        env.reset_tensor()
    cur = 0
    for t in range(25):
        a = torch.randint(0, 5, (n,), generator=g, device=dev)
        _, rw, dw = whole.step_tensor(a)
        parts = [h.step_tensor(a[q * (n // 2):(q + 1) * (n // 2)].contiguous()) for q, h in enumerate(halves)]
        mw = torch.zeros(3, dtype=torch.float64, device=dev)
        mab = torch.zeros((2, 3), dtype=torch.float64, device=dev)
        _lib.check(L.ptg_vecnorm_moments(whole._h, p(rw), p(ret["w"]), 0.99, p(st["w"][cur]), p(mw), stream))
        for q, (h, key) in enumerate(zip(halves, "ab")):
            _lib.check(L.ptg_vecnorm_moments(h._h, p(parts[q][1]), p(ret[key]), 0.99, p(st[key][cur]), p(mab[q]), stream))
        _lib.check(L.ptg_vecnorm_apply(whole._h, p(rw), p(dw), p(ret["w"]), p(st["w"][cur]), p(st["w"][cur ^ 1]), p(mw), 1, 1,
                                       1e-8, 10.0, p(out["w"]), stream))
        for q, (h, key) in enumerate(zip(halves, "ab")):
            _lib.check(L.ptg_vecnorm_apply(h._h, p(parts[q][1]), p(parts[q][2]), p(ret[key]), p(st[key][cur]),
                                           p(st[key][cur ^ 1]), p(mab), 2, 1, 1e-8, 10.0, p(out[key]), stream))
        cur ^= 1
        sw, sa, sb = (st[k][cur].cpu().numpy() for k in "wab")
        assert np.array_equal(sa, sb)                                    # both "ranks" hold identical statistics
        assert np.allclose(sa[:3], sw[:3], rtol=1e-11, atol=1e-13)
        assert np.allclose(torch.cat([out["a"], out["b"]]).cpu().numpy(), out["w"].cpu().numpy(), rtol=1e-6, atol=1e-9)
    for env in [whole] + halves:
        env.close()


@pytest.mark.parametrize("design", ["mod", "raw"])
@pytest.mark.parametrize("n_envs", [7, 256, 1000])
def test_flat_layout_equals_dict_layout_plus_feature_kernel(design, n_envs):
    """obs_layout="flat": the step kernel writes the CombinedExtractor rows itself.  Against the key-major env +
    ptg_features on the same actions: identical rows, rewards, dones, states -- through resets, an episode end
    (auto-reset + terminal observation), the roll-out kernel and the numpy API."""
    import torch
    from rl_ptg_b200.vec_env import PtGVecEnv
    from rl_ptg_b200.vec_normalize import features_tensor
    kw = dict(synthetic_kwargs(dict(scenario=3, operation="OP2", raw_modified=design)))
    kw["eps_sim_steps"] = 40                              # short episodes: the auto-reset path runs several times
    ed = PtGVecEnv(kw, n_envs, seed=5)
    ef = PtGVecEnv(kw, n_envs, seed=5, obs_layout="flat")
    F = ef.feature_dim
    assert F == (40 if design == "mod" else 31) and ef.obs_elems >= n_envs * F
    ed.reset_tensor(); ef.reset_tensor()
    assert torch.equal(features_tensor(ed), ef.features_view())
    assert features_tensor(ef).data_ptr() == ef._obs.data_ptr()          # zero-copy for the flat env
    g = torch.Generator(device=ed.device); g.manual_seed(9)
    for t in range(90):
        a = torch.randint(0, 5, (n_envs,), generator=g, device=ed.device)
        _, r1, d1 = ed.step_tensor(a)
        _, r2, d2 = ef.step_tensor(a)
        assert torch.equal(r1, r2) and torch.equal(d1, d2), f"step {t}"
        assert torch.equal(features_tensor(ed), ef.features_view()), f"rows differ at step {t}"
        if bool(d1.any()):
            done = d1.bool()
            assert torch.equal(features_tensor(ed, ed._term_obs)[done], ef.features_view(ef._term_obs)[done])
    s1, s2 = ed.get_state(), ef.get_state()
    for f in ("meth_state", "i", "j", "k", "hot_cold", "draws", "episode_count", "act_ep_h"):
        assert np.array_equal(s1[f], s2[f]), f
    # roll-out kernel
    acts = torch.randint(0, 5, (12, n_envs), generator=g, device=ed.device)
    ro_d, ro_f = ed.rollout_tensor(acts), ef.rollout_tensor(acts)
    assert torch.equal(ro_d["reward"], ro_f["reward"]) and torch.equal(ro_d["done"], ro_f["done"])
    for t in range(12):
        assert torch.equal(features_tensor(ed, ro_d["obs"][t]), ef.features_view(ro_f["obs"][t])), f"rollout step {t}"
    # masked reset + numpy API (dict of arrays, METH_STATUS back to Discrete values)
    mask = (np.arange(n_envs) % 3 == 0).astype(np.uint8)
    ed.reset_tensor(mask=mask); ef.reset_tensor(mask=mask)
    assert torch.equal(features_tensor(ed), ef.features_view())
    a = np.random.default_rng(0).integers(0, 5, n_envs)
    o1, r1, d1, _ = ed.step(a)
    o2, r2, d2, _ = ef.step(a)
    assert set(o1) == set(o2)
    for key in o1:
        assert np.array_equal(np.asarray(o1[key]).reshape(n_envs, -1), np.asarray(o2[key]).reshape(n_envs, -1)), key
    assert np.array_equal(r1, r2) and np.array_equal(d1, d2)
    ed.close(); ef.close()


def test_flat_layout_limits_fail_loudly():
    from rl_ptg_b200._lib import PtgError
    from rl_ptg_b200.vec_env import PtGVecEnv
    with pytest.raises(PtgError, match="UNSUPPORTED"):
        PtGVecEnv(synthetic_kwargs(dict(scenario=2, operation="OP2", price_ahead=6)), 8, obs_layout="flat")
    with pytest.raises(PtgError, match="UNSUPPORTED"):
        PtGVecEnv(synthetic_kwargs(dict(scenario=2, operation="OP2")), 8, train_or_eval="eval", obs_layout="flat")
    with pytest.raises(ValueError):
        PtGVecEnv(synthetic_kwargs(dict(scenario=2, operation="OP2")), 8, obs_layout="rows")


def test_calculate_optimum_on_device_is_bit_identical():
    """calculate_optimum (src/rl_opt.py:26-152) with the per-hour work in a CUDA kernel --
    all 24 columns bit-identical to the host version (which tests/test_preprocessing.py pins to vectors recorded from
    the unmodified reference) and to those vectors directly; real data, three scenario / operation-point configs,
    full training series (36 552 h)."""
    import json
    import os
    import rl_ptg_b200 as ptg
    from helpers import GOLDEN_DIR, REF_DATA
    from rl_ptg_b200.config import STATS_NAMES
    from rl_ptg_b200.preprocessing import calculate_optimum, calculate_optimum_cuda
    with np.load(os.path.join(GOLDEN_DIR, "topt_reference.npz")) as z:
        meta = json.loads(str(z["meta"]))
        want = {m["name"]: z[m["name"]] for m in meta}
    for m in meta:
        E = ptg.EnvConfiguration(**m["overrides"])
        price, _ = ptg.load_data_npz(REF_DATA, E)
        for split in (m["split"], "train"):
            args = (price[f"el_price_{split}"], price[f"gas_price_{split}"], price[f"eua_price_{split}"], "set",
                    list(STATS_NAMES), E)
            host = calculate_optimum(*args)
            dev = calculate_optimum_cuda(*args)
            for k in STATS_NAMES:
                assert np.array_equal(host[k], dev[k]), (m["name"], split, k)
        got = np.stack([calculate_optimum_cuda(price[f"el_price_{m['split']}"], price[f"gas_price_{m['split']}"],
                                               price[f"eua_price_{m['split']}"], "set", list(STATS_NAMES), E)[k]
                        for k in STATS_NAMES], axis=1)
        assert np.array_equal(got, want[m["name"]])


def test_eval_callback_keeps_the_best_policy():
    from rl_ptg_b200.ppo import EvalCallback, PPO, reference_hyper_kwargs
    from rl_ptg_b200.vec_env import PtGVecEnv
    env = _env(1024, obs_layout="flat")
    val = PtGVecEnv(synthetic_kwargs(dict(scenario=2, operation="OP2"), split="val"), 8, seed=605, obs_layout="flat")
    hyper = reference_hyper_kwargs()
    hyper.update(n_steps=16, batch_size=4096, n_epochs=2, learning_rate=3e-4, seed=0)
    model = PPO(env, **hyper)
    cb = EvalCallback(val, n_eval_steps=200, eval_freq=2 * 16 * 1024)
    model.learn(6 * 16 * 1024, callback=cb)
    assert len(cb.evaluations) == 3 and cb.best_state_dict is not None
    assert cb.best_mean_reward == max(e["mean_cum_reward"] for e in cb.evaluations)
    assert set(cb.best_state_dict) == set(model.policy.state_dict())
    env.close(); val.close()
