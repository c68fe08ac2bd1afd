"""Host preprocessing (config -> data -> e_r_b/g_e/eps_ind/kwargs) against values recorded from the reference."""
import numpy as np
import pytest

import rl_ptg_b200 as ptg
from helpers import GOLDEN_CASES, golden_kwargs, load_golden, synthetic_kwargs


@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_kwargs_match_reference(case):
    g, kw = load_golden(case), golden_kwargs(case)
    assert kw["rew_l_b"] == g["rew_l_b"] and kw["rew_u_b"] == g["rew_u_b"]
    assert np.array_equal(np.asarray(kw["reward_level"]), g["reward_level"])
    assert kw["max_h2_volumeflow"] == g["max_h2_volumeflow"]
    assert kw["eps_sim_steps"] == int(g["eps_sim_steps"]) and kw["n_eps_loops"] == int(g["n_eps_loops"])
    if kw["eps_ind"] is None:
        assert g["eps_ind"].size == 0
    else:
        assert np.array_equal(kw["eps_ind"], g["eps_ind"])
    assert np.array_equal(kw["e_r_b"].sum(axis=2), g["e_r_b_sum"])
    assert np.array_equal(kw["g_e"].sum(axis=2), g["g_e_sum"])
    assert np.array_equal(kw["e_r_b"][:, :, ::997], g["e_r_b_probe"])


def test_topt_known_answers():
    # T-OPT BS2/OP2 train/val/test (printed by src/rl_opt.py:147-148)
    E = ptg.EnvConfiguration()
    from helpers import REF_DATA
    price, _ = ptg.load_data_npz(REF_DATA, E)
    for split, want in (("train", 1200689.31), ("val", 50489.62), ("test", 27043.46)):
        s = ptg.calculate_optimum(price[f"el_price_{split}"], price[f"gas_price_{split}"],
                                  price[f"eua_price_{split}"], split, E.stats_names, E)
        assert round(s["Meth_cum_reward_stats"][-E.price_ahead], 2) == want


def test_appendix_b_known_answers():
    kw = golden_kwargs("bs2_op2_mod")
    assert kw["eps_ind"][:12].tolist() == [33, 3, 38, 0, 8, 23, 9, 5, 35, 36, 2, 4]
    assert len(kw["eps_ind"]) == 2460 and kw["eps_sim_steps"] == 5328
    assert kw["e_r_b"].shape == (3, 13, 36539) and kw["g_e"].shape == (2, 2, 1522)
    assert kw["rew_l_b"] == -685.2843133116897 and kw["rew_u_b"] == 1372.5540663690715


def test_config_validation():
    with pytest.raises(ValueError):
        ptg.EnvConfiguration(scenario=4)
    with pytest.raises(ValueError):
        ptg.EnvConfiguration(operation="OP3")
    with pytest.raises(ValueError):
        ptg.EnvConfiguration(raw_modified="both")
    with pytest.raises(KeyError):
        ptg.EnvConfiguration(not_a_knob=1)
    with pytest.raises(ValueError):
        ptg.load_data_npz.__globals__["_finish_price_dict"](
            {**{f"{k}_price_{s}": np.ones(24 * 50 if k == "el" else 50) for k in ("el", "gas", "eua")
                for s in ("train", "val", "test")}}, ptg.EnvConfiguration())   # 44 days not divisible by 37


def test_synthetic_shapes():
    from rl_ptg_b200 import _abi
    kw = synthetic_kwargs()
    assert kw["e_r_b"].shape == (3, 13, 36539) and kw["g_e"].shape == (2, 2, 1522)
    assert kw["cooldown"].shape == (45001, 7) and sum(kw[k].shape[0] for k in _abi.DATASET_NAMES) == 118760
    assert set(np.unique(kw["e_r_b"][2])) <= {-1.0, 0.0, 1.0}


def test_calculate_optimum_matches_reference_all_columns():
    """All 24 columns of calculate_optimum (src/rl_opt.py:26-152) bit for bit against vectors recorded from the
    unmodified reference (tests/golden/gen_golden_topt.py) -- including the reference's quirk that the revenue /
    cost constituents always carry the full-load values."""
    import json
    import os
    import numpy as np
    import rl_ptg_b200 as ptg
    from helpers import GOLDEN_DIR, REF_DATA
    from rl_ptg_b200.config import STATS_NAMES
    from rl_ptg_b200.preprocessing import calculate_optimum
    with np.load(os.path.join(GOLDEN_DIR, "topt_reference.npz")) as z:
        meta = json.loads(str(z["meta"]))
        want = {m["name"]: z[m["name"]] for m in meta}
    for m in meta:
        E = ptg.EnvConfiguration(**m["overrides"])
        price, _ = ptg.load_data_npz(REF_DATA, E)
        s = m["split"]
        d = calculate_optimum(price[f"el_price_{s}"], price[f"gas_price_{s}"], price[f"eua_price_{s}"], "Test_set",
                              list(STATS_NAMES), E)
        got = np.stack([d[k] for k in STATS_NAMES], axis=1)
        assert list(STATS_NAMES) == m["stats_names"]
        assert np.array_equal(got, want[m["name"]]), m["name"]
