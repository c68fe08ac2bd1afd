"""Configuration objects of the batched PtG environment.

Host-side mirror of the reference's YAML -> object config layer for the hot path:
``EnvConfiguration`` follows ``src/rl_config_env.py:12-49`` (same attribute names, same derived values
``max_h2_volumeflow`` / ``h2_u_b`` / ``ch4_u_b`` / ``h2_res_u_b`` / ``h2o_u_b`` / ``datafile_path2..18`` /
``stats_names``), ``TrainConfiguration`` carries only the fields that reach the environment path
(``config/config_train.yaml``: ``parallel``, ``train_or_eval``, ``train_steps``, ``r_seed_train/test``) and
``AgentConfiguration`` only ``n_envs`` and ``rl_alg_hyp["action_type"]`` (``config/config_agent.yaml:11,16``).
Unlike the reference the YAML path is an argument (the reference opens ``config/config_env.yaml`` relative to
the cwd, ``src/rl_config_env.py:15``) and individual knobs can be overridden by keyword.
"""
from __future__ import annotations

import os
from typing import Any

import yaml

_DEFAULT_ENV_YAML = os.path.join(os.path.dirname(__file__), "config", "config_env.yaml")

OP_DATASETS = (  # dict key -> datafile_path index (reference src/rl_utils.py:104-110)
    ("startup_cold", 2), ("startup_hot", 3), ("cooldown", 4), ("standby_down", 5), ("standby_up", 6),
    ("op1_start_p", 7), ("op2_start_f", 8), ("op3_p_f", 9), ("op4_p_f_p_5", 10), ("op5_p_f_p_10", 11),
    ("op6_p_f_p_15", 12), ("op7_p_f_p_22", 13), ("op8_f_p", 14), ("op9_f_p_f_5", 15), ("op10_f_p_f_10", 16),
    ("op11_f_p_f_15", 17), ("op12_f_p_f_20", 18),
)

STATS_NAMES = (
    "steps_stats", "el_price_stats", "gas_price_stats", "eua_price_stats", "Meth_State_stats",
    "Meth_Action_stats", "Meth_Hot_Cold_stats", "Meth_T_cat_stats", "Meth_H2_flow_stats",
    "Meth_CH4_flow_stats", "Meth_H2O_flow_stats", "Meth_el_heating_stats", "Meth_ch4_revenues_stats",
    "Meth_steam_revenues_stats", "Meth_o2_revenues_stats", "Meth_eua_revenues_stats",
    "Meth_chp_revenues_stats", "Meth_elec_costs_heating_stats", "Meth_elec_costs_electrolyzer_stats",
    "Meth_water_costs_stats", "Meth_reward_stats", "Meth_cum_reward_stats", "pot_reward_stats",
    "part_full_stats",
)


class EnvConfiguration:
    """All environment knobs; attribute-compatible with the reference's ``EnvConfiguration``."""

    def __init__(self, path: str | None = None, **overrides: Any):
        with open(path or _DEFAULT_ENV_YAML, "r") as fh:
            cfg = yaml.safe_load(fh)
        unknown = set(overrides) - set(cfg)
        if unknown:
            raise KeyError(f"unknown config_env knobs: {sorted(unknown)}")
        cfg.update(overrides)
        self.__dict__.update(cfg)

        if self.scenario not in (1, 2, 3):
            raise ValueError(f"scenario ({self.scenario}) must be one of [1, 2, 3]")
        if self.raw_modified not in ("raw", "mod"):
            raise ValueError(f"raw_modified ({self.raw_modified}) must be 'raw' or 'mod'")
        if self.operation not in ("OP1", "OP2"):
            raise ValueError(f"operation ({self.operation}) must be 'OP1' or 'OP2'")
        self.train_len_d = None  # filled by load_data()

        base = self.datafile_path["path"] + self.operation
        for _, num in OP_DATASETS:
            setattr(self, f"datafile_path{num}", f"{base}/{self.datafile_path['datafile'][f'datafile_path{num}']}")

        self.meth_stats_load = self.meth_stats_load[self.operation]
        full = 2  # [off, partial_load, full_load]
        self.max_h2_volumeflow = self.convert_mol_to_Nm3 * self.meth_stats_load["Meth_H2_flow"][full]
        self.h2_u_b = self.meth_stats_load["Meth_H2_flow"][full]
        self.ch4_u_b = self.meth_stats_load["Meth_CH4_flow"][full]
        self.h2_res_u_b = self.meth_stats_load["Meth_H2_res_flow"][full]
        self.h2o_u_b = self.meth_stats_load["Meth_H2O_flow"][full]
        self.stats_names = list(STATS_NAMES)


class TrainConfiguration:
    """The four training-config fields that reach the env path (config/config_train.yaml:24-36)."""

    def __init__(self, parallel: str = "Singleprocessing", train_or_eval: str = "train",
                 train_steps: int = 1_500_000, seed_train: int = 3654, seed_test: int = 605,
                 path: str | None = None, eval_trials: int = 5):
        if parallel not in ("Singleprocessing", "Multiprocessing"):
            raise ValueError('parallel must be "Singleprocessing" or "Multiprocessing"')
        if train_or_eval not in ("train", "eval"):
            raise ValueError('train_or_eval must be "train" or "eval"')
        self.parallel = parallel
        self.train_or_eval = train_or_eval
        self.train_steps = train_steps
        self.seed_train = seed_train
        self.seed_test = seed_test
        self.path = path
        self.eval_trials = eval_trials


class AgentConfiguration:
    """Only what the env path reads: ``n_envs`` and the algorithm's ``action_type``."""

    def __init__(self, n_envs: int = 6, action_type: str = "discrete"):
        if action_type not in ("discrete", "continuous"):
            raise ValueError("action_type must be 'discrete' or 'continuous'")
        self.n_envs = n_envs
        self.rl_alg_hyp = {"action_type": action_type}
