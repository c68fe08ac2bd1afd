"""CPU-side checks of the product: the C-ABI library loads and exports every symbol include/ptg_b200.h declares,
struct layouts agree with the header, host twins of the device RNG are bit-identical to numpy, config marshalling
validates like the reference's asserts, and nothing under rl_ptg_b200/ touches the oracle."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from helpers import ROOT_DIR, synthetic_kwargs
from rl_ptg_b200 import _abi, _lib, spaces


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT_DIR, "include", "ptg_b200.h")).read()
    declared = set(re.findall(r"\b(ptg_[a-z0-9_]+)\s*\(", hdr))
    L = _lib.load()
    assert declared, "no declarations parsed"
    for name in sorted(declared):
        assert hasattr(L, name), f"libptg_b200.so does not export {name}"
    assert declared == set(_lib.EXPORTS)
    assert L.ptg_abi_version() == _abi.PTG_ABI_VERSION


def test_struct_layouts_match_header():
    # PtgConfig: 36 int32 then 43 doubles, no padding surprises
    assert C.sizeof(_abi.PtgConfig) == 36 * 4 + 43 * 8
    assert _abi.PtgConfig.noise.offset == 144
    assert C.sizeof(_abi.PtgTables) == 17 * 8 * 2 + 6 * 8
    assert C.sizeof(_abi.PtgIO) == 9 * 8
    assert C.sizeof(_abi.PtgObsKey) == 40
    assert C.sizeof(_abi.PtgEpisodeStats) == 64
    assert C.sizeof(_abi.PtgStateSoA) == 18 * 8


def test_stats_combine_host():
    L = _lib.load()
    arr = (_abi.PtgEpisodeStats * 3)()
    vals = [(2, 10.0, 60.0, 20, 4.0, 6.0, 100), (0, 0.0, 0.0, 0, np.inf, -np.inf, 50), (1, -3.0, 9.0, 7, -3.0, -3.0, 25)]
    for a, v in zip(arr, vals):
        a.count, a.sum_return, a.sum_return_sq, a.sum_length, a.min_return, a.max_return, a.total_steps = v
    out = _abi.PtgEpisodeStats()
    L.ptg_stats_combine(arr, 3, C.byref(out))
    assert (out.count, out.sum_return, out.sum_return_sq, out.sum_length) == (3, 7.0, 69.0, 27)
    assert (out.min_return, out.max_return, out.total_steps) == (-3.0, 6.0, 175)


@pytest.mark.parametrize("seed", [0, 1, 3654, 605, 2 ** 32 - 1, 2 ** 32, 2 ** 40 + 17])
def test_host_rng_twin_is_numpy_exact(seed):
    """Same source file as the device generator (csrc/ptg_rng.cuh): SeedSequence -> PCG64 -> ziggurat."""
    L = _lib.load()
    st = (C.c_uint64 * 4)()
    assert L.ptg_host_seed_state(seed, st) == 0
    ref = np.random.PCG64(seed).state["state"]
    assert (int(st[0]) << 64) | int(st[1]) == ref["state"] and (int(st[2]) << 64) | int(st[3]) == ref["inc"]
    n = 200_000
    out = np.empty(n)
    assert L.ptg_host_standard_normal(seed, n, out.ctypes.data) == 0
    assert np.array_equal(out, np.random.default_rng(seed).standard_normal(n))


def test_config_marshalling_validates_like_the_reference():
    kw = dict(synthetic_kwargs())
    cfg = _abi.config_from_kwargs(kw, "train")
    assert cfg.price_ahead == 13 and cfg.sim_step == 600 and cfg.eps_sim_steps == 5328 and cfg.raw_modified == 1
    assert cfg.reward_level == float(np.asarray(kw["reward_level"])[0])
    with pytest.raises(ValueError):
        _abi.config_from_kwargs(kw, "test")                       # ptg_gym_env.py:46
    with pytest.raises(ValueError):
        _abi.config_from_kwargs({**kw, "action_type": "box"})     # :158
    with pytest.raises(ValueError):
        _abi.config_from_kwargs({**kw, "raw_modified": "x"})      # :204
    with pytest.raises(ValueError):
        _abi.config_from_kwargs({**kw, "ptg_standby": 3})
    with pytest.raises(ValueError):
        _abi.tables_from_kwargs({**kw, "cooldown": np.zeros((5, 6))}, 13)
    t, keep = _abi.tables_from_kwargs(kw, 13)
    assert t.n_hours == 36539 and t.n_days == 1522 and t.n_eps_ind == len(kw["eps_ind"])
    assert _abi.tables_from_kwargs({**kw, "eps_ind": None}, 13)[0].n_eps_ind == 0


def test_spaces_match_reference_definition():
    obs = spaces.observation_space("mod", 13)
    assert list(obs.spaces.keys()) == ["Pot_Reward", "Part_Full", "METH_STATUS", "T_CAT", "H2_in_MolarFlow",
                                       "CH4_syn_MolarFlow", "H2_res_MolarFlow", "H2O_DE_MassFlow", "Elec_Heating",
                                       "Temp_hour_enc_sin", "Temp_hour_enc_cos"]
    assert obs["Pot_Reward"].shape == (13,) and obs["Part_Full"].low.min() == -1 and obs["METH_STATUS"].n == 6
    raw = spaces.observation_space("raw", 13)
    assert list(raw.spaces.keys())[:3] == ["Elec_Price", "Gas_Price", "EUA_Price"] and raw["Gas_Price"].shape == (2,)
    assert [k for k, _, _ in _abi.obs_keys("raw", 13)] == list(raw.spaces.keys())
    assert spaces.action_space("discrete").n == 5
    box = spaces.action_space("continuous")
    assert box.shape == (1,) and box.dtype == np.float32 and box.low[0] == -1 and box.high[0] == 1
    with pytest.raises(ValueError):
        spaces.action_space("other")


def test_product_never_touches_the_oracle():
    """The oracle is test infrastructure: no file of the product package may import, load or mention it."""
    pkg = os.path.join(ROOT_DIR, "rl_ptg_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.lower(), f"{f} references the oracle"


def test_vec_env_requires_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from rl_ptg_b200.vec_env import PtGVecEnv
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        PtGVecEnv(synthetic_kwargs(), 4)


def test_lazy_infos_behaves_like_a_list_of_dicts():
    from rl_ptg_b200.vec_env import LazyInfos
    made = []

    def make(e):
        made.append(e)
        return {"env": e}

    infos = LazyInfos(1_000_000, None, make)
    assert len(infos) == 1_000_000 and made == []
    assert infos[5] == {"env": 5} and infos[-1] == {"env": 999_999}
    infos[5]["extra"] = 1
    assert infos[5]["extra"] == 1 and made == [5, 999_999]          # identity is stable, nothing else materialised
    assert [d["env"] for d in infos[10:13]] == [10, 11, 12]
    with pytest.raises(IndexError):
        infos[1_000_000]
    empty = LazyInfos(3)
    assert list(empty) == [{}, {}, {}]
