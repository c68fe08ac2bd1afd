"""One-off host pre-processing that feeds the env path (reference L1).

* ``calculate_optimum`` -- per-hour steady-state reward of partial/full load, the "potential reward" and the
  load identifier (reference ``src/rl_opt.py:26-152``).  The reference runs a Python double loop over
  36 552 h x 2 loads; here the load-dependent scalars are evaluated once with Python floats (same operation
  order as ``rl_opt.py:53-89``) and only the price-dependent products are vectorised over the hours, which
  yields bit-identical fp64 values.
* ``Preprocessing`` -- ``e_r_b`` / ``g_e`` price tensors (``src/rl_utils.py:243-281``), the episode schedule
  ``eps_ind`` (``:283-335``) and ``dict_env_kwargs`` (``:337-405``): the flat ~95-key dict that IS the
  constructor ABI of the environment.
"""
from __future__ import annotations

import math

import numpy as np

from .config import AgentConfiguration, EnvConfiguration, TrainConfiguration

_SCALAR_KEYS = (
    "noise", "eps_len_d", "sim_step", "time_step_op", "price_ahead", "scenario",
    "convert_mol_to_Nm3", "H_u_CH4", "H_u_H2", "dt_water", "cp_water", "rho_water",
    "Molar_mass_CO2", "Molar_mass_H2O", "h_H2O_evap", "eeg_el_price", "heat_price",
    "o2_price", "water_price", "min_load_electrolyzer", "max_h2_volumeflow", "eta_CHP",
    "t_cat_standby", "t_cat_startup_cold", "t_cat_startup_hot", "time1_start_p_f",
    "time2_start_f_p", "time_p_f", "time_f_p", "time1_p_f_p", "time2_p_f_p",
    "time23_p_f_p", "time3_p_f_p", "time34_p_f_p", "time4_p_f_p", "time45_p_f_p",
    "time5_p_f_p", "time1_f_p_f", "time2_f_p_f", "time23_f_p_f", "time3_f_p_f",
    "time34_f_p_f", "time4_f_p_f", "time45_f_p_f", "time5_f_p_f", "i_fully_developed",
    "j_fully_developed", "el_l_b", "el_u_b", "gas_l_b", "gas_u_b", "eua_l_b", "eua_u_b",
    "T_l_b", "T_u_b", "h2_l_b", "h2_u_b", "ch4_l_b", "ch4_u_b", "h2_res_l_b", "h2_res_u_b",
    "h2o_l_b", "h2o_u_b", "heat_l_b", "heat_u_b", "raw_modified",
)


def electrolyzer_efficiency(load: float, min_load: float) -> float:
    """LHV efficiency of the PEM electrolyzer vs. load (rl_opt.py:82-88 == ptg_gym_env.py:312-317)."""
    if load < min_load:
        return 0.02
    return (0.598 - 0.325 * load ** 2 + 0.218 * load ** 3 + 0.01 * load ** (-1)
            - 1.68 * 10 ** (-3) * load ** (-2) + 2.51 * 10 ** (-5) * load ** (-3))


def calculate_optimum(el_price_data, gas_price_data, eua_price_data, data_name: str, stats_names,
                      EnvConfig: EnvConfiguration, verbose: bool = False) -> dict:
    """Theoretical optimum ignoring plant dynamics; returns the reference's 24-column stats dict.

    Unlike the reference (which re-reads ``config/config_env.yaml`` from the cwd, rl_opt.py:37) the config
    object is an argument.
    """
    C = EnvConfig
    ms = C.meth_stats_load
    el = np.asarray(el_price_data, dtype=np.float64)
    gas = np.asarray(gas_price_data, dtype=np.float64)
    eua = np.asarray(eua_price_data, dtype=np.float64)
    n = len(el)
    t_day = np.arange(n) // 24
    t_day = np.where(t_day == len(gas), t_day - 1, t_day)      # rl_opt.py:52
    gas_h, eua_h = gas[t_day], eua[t_day]
    b_s3 = 1 if C.scenario == 3 else 0

    rew_l, parts = [], []
    for l in (1, 2):  # partial, full
        ch4_volumeflow = ms["Meth_CH4_flow"][l] * C.convert_mol_to_Nm3
        h2_res_volumeflow = ms["Meth_H2_res_flow"][l] * C.convert_mol_to_Nm3
        Q_ch4 = ch4_volumeflow * C.H_u_CH4 * 1000
        Q_h2_res = h2_res_volumeflow * C.H_u_H2 * 1000
        ch4_revenues = (Q_ch4 + Q_h2_res) * gas_h
        power_chp = Q_ch4 * C.eta_CHP * b_s3
        Q_chp = Q_ch4 * (1 - C.eta_CHP) * b_s3
        chp_revenues = power_chp * C.eeg_el_price
        Q_steam = ms["Meth_H2O_flow"][l] * (C.dt_water * C.cp_water + C.h_H2O_evap) / 3600
        steam_revenues = (Q_steam + Q_chp) * C.heat_price
        h2_volumeflow = ms["Meth_H2_flow"][l] * C.convert_mol_to_Nm3
        o2_volumeflow = 1 / 2 * h2_volumeflow * 3600
        o2_revenues = o2_volumeflow * C.o2_price
        Meth_CO2_mass_flow = ms["Meth_CH4_flow"][l] * C.Molar_mass_CO2 / 1000
        eua_revenues = Meth_CO2_mass_flow / 1000 * 3600 * eua_h * 100
        elec_costs_heating = ms["Meth_el_heating"][l] / 1000 * el
        load_elec = h2_volumeflow / C.max_h2_volumeflow
        eta = electrolyzer_efficiency(load_elec, C.min_load_electrolyzer)
        elec_costs_electrolyzer = h2_volumeflow * C.H_u_H2 * 1000 / eta * el
        elec_costs = elec_costs_heating + elec_costs_electrolyzer
        water_elec = ms["Meth_H2_flow"][l] * C.Molar_mass_H2O / 1000 * 3600
        water_costs = (ms["Meth_H2O_flow"][l] + water_elec) / C.rho_water * C.water_price
        rew_l.append(ch4_revenues + chp_revenues + steam_revenues + eua_revenues + o2_revenues - elec_costs
                     - water_costs)
        parts.append((ch4_revenues, steam_revenues, o2_revenues, eua_revenues, chp_revenues,
                      -elec_costs_heating, -elec_costs_electrolyzer, -water_costs))

    index = (rew_l[1] > rew_l[0]).astype(np.int64)     # list.index(max(...)): ties -> partial load
    rew = np.where(index == 1, rew_l[1], rew_l[0])
    on = rew > 0

    stats = np.zeros((n, len(stats_names)))
    stats[:, 0] = np.arange(n)
    stats[:, 1], stats[:, 2], stats[:, 3] = el, gas_h, eua_h
    level = np.where(on, index + 1, 0)
    for col, key in enumerate(("Meth_State", "Meth_Action", "Meth_Hot_Cold", "Meth_T_cat", "Meth_H2_flow",
                               "Meth_CH4_flow", "Meth_H2O_flow", "Meth_el_heating"), start=4):
        stats[:, col] = np.asarray(ms[key], dtype=np.float64)[level]
    for col in range(8):
        # reference quirk (rl_opt.py:113-124): the constituent columns are filled from the loop variables AFTER the
        # partial/full loop, i.e. they always hold the FULL-load values, also where partial load was the better choice
        b = np.broadcast_to(np.asarray(parts[1][col], dtype=np.float64), (n,))
        stats[:, 12 + col] = np.where(on, b, 0.0)
    stats[:, 20] = rew
    stats[:, 21] = np.cumsum(np.where(on, rew, 0.0))
    stats[:, 23] = np.where(on, index, -1)

    stats_dict_opt = {name: stats[:, m] for m, name in enumerate(stats_names)}
    if verbose and data_name != "reward_Level":
        print("    > ", data_name, ": Cumulative reward - theoretical optimum T-OPT = ",
              round(stats_dict_opt["Meth_cum_reward_stats"][-C.price_ahead], 2))
    return stats_dict_opt


class Preprocessing:
    """Builds everything ``PTGEnv``/``PtGVecEnv`` takes as ``dict_input`` (reference class of the same name)."""

    def __init__(self, dict_price_data, dict_op_data, AgentConfig: AgentConfiguration, EnvConfig: EnvConfiguration,
                 TrainConfig: TrainConfiguration, verbose: bool = False):
        self.AgentConfig, self.EnvConfig, self.TrainConfig = AgentConfig, EnvConfig, TrainConfig
        self.dict_price_data, self.dict_op_data = dict_price_data, dict_op_data
        self.verbose = verbose
        self.overhead_factor = 10       # rl_utils.py:187
        self.preprocessing_rew()
        self.preprocessing_array()
        self.define_episodes()

    # -- potential reward / load identifier (rl_utils.py:202-240) ---------------------------------------------
    def preprocessing_rew(self):
        d, E = self.dict_price_data, self.EnvConfig
        opt = {split: calculate_optimum(d[f"el_price_{split}"], d[f"gas_price_{split}"], d[f"eua_price_{split}"],
                                        name, E.stats_names, E, self.verbose)
               for split, name in (("train", "Training"), ("val", "Validation"), ("test", "Test"))}
        level = calculate_optimum(d["el_price_reward_level"], d["gas_price_reward_level"],
                                  d["eua_price_reward_level"], "reward_Level", E.stats_names, E)
        self.stats_dict_opt = opt
        self.dict_pot_r_b = {}
        for split in ("train", "val", "test"):
            self.dict_pot_r_b[f"pot_rew_{split}"] = opt[split]["Meth_reward_stats"]
            self.dict_pot_r_b[f"part_full_b_{split}"] = opt[split]["part_full_stats"]
        self.r_level = level["Meth_reward_stats"]

    # -- sliding windows (rl_utils.py:243-281) -----------------------------------------------------------------
    def preprocessing_array(self):
        pa = self.EnvConfig.price_ahead
        for split in ("train", "val", "test"):
            series = (self.dict_price_data[f"el_price_{split}"], self.dict_pot_r_b[f"pot_rew_{split}"],
                      self.dict_pot_r_b[f"part_full_b_{split}"])
            n = series[0].shape[0] - pa
            e_r_b = np.zeros((3, pa, n))
            for c, s in enumerate(series):
                e_r_b[c] = np.lib.stride_tricks.sliding_window_view(s, n)[:pa]     # [a, t] = s[t + a]
            gas, eua = self.dict_price_data[f"gas_price_{split}"], self.dict_price_data[f"eua_price_{split}"]
            g_e = np.zeros((2, 2, gas.shape[0] - 1))
            g_e[0, 0], g_e[0, 1] = gas[:-1], gas[1:]
            g_e[1, 0], g_e[1, 1] = eua[:-1], eua[1:]
            setattr(self, f"e_r_b_{split}", e_r_b)
            setattr(self, f"g_e_{split}", g_e)

    # -- episodes (rl_utils.py:283-335) -------------------------------------------------------------------------
    def define_episodes(self):
        E, T, d = self.EnvConfig, self.TrainConfig, self.dict_price_data
        val_len_d = len(d["gas_price_val"]) - 1
        test_len_d = len(d["gas_price_test"]) - 1
        self.n_eps = int(E.train_len_d / E.eps_len_d)
        self.eps_len = 24 * 3600 * E.eps_len_d
        self.eps_sim_steps_train = int(self.eps_len / E.sim_step)
        self.eps_sim_steps_val = int(24 * 3600 * val_len_d / E.sim_step)
        self.eps_sim_steps_test = int(24 * 3600 * test_len_d / E.sim_step)
        self.num_loops = T.train_steps / (self.eps_sim_steps_train * self.n_eps)
        self.rand_eps_ind()
        self.n_eps_loops = self.n_eps * int(self.num_loops)

    def rand_eps_ind(self):
        """Episode order: independent shuffles of 0..n_eps-1 under the legacy global seed (rl_utils.py:315-335)."""
        rs = np.random.RandomState(self.TrainConfig.seed_train)   # == np.random.seed(seed) + np.random.shuffle
        E = self.EnvConfig
        if E.train_len_d == E.eps_len_d:
            self.eps_ind = np.zeros(self.n_eps * int(self.num_loops) * self.overhead_factor)
            return
        loops = 1 if self.num_loops < 1 else int(self.num_loops)
        random_ep = np.tile(np.linspace(0, self.n_eps - 1, self.n_eps), (loops * self.overhead_factor, 1))
        for row in random_ep:
            rs.shuffle(row)
        self.eps_ind = random_ep.reshape(-1).astype(int)

    # -- constructor ABI of the env (rl_utils.py:337-405) ---------------------------------------------------------
    def dict_env_kwargs(self, type: str = "train") -> dict:
        E = self.EnvConfig
        kw = {f"ptg_{k}": E.ptg_state_space[k] for k in ("standby", "cooldown", "startup", "partial_load",
                                                        "full_load")}
        kw.update({k: getattr(E, k) for k in _SCALAR_KEYS})
        kw.update(parallel=self.TrainConfig.parallel, n_eps_loops=self.n_eps_loops, reward_level=self.r_level,
                  action_type=self.AgentConfig.rl_alg_hyp["action_type"])
        kw.update(self.dict_op_data)
        if type not in ("train", "val", "test"):
            raise ValueError(f'Invalid type: {type}. Must be "train", "val", or "test".')
        train_pot = self.e_r_b_train[1, 0, :]
        kw.update(e_r_b=getattr(self, f"e_r_b_{type}"), g_e=getattr(self, f"g_e_{type}"),
                  eps_sim_steps=getattr(self, f"eps_sim_steps_{type}"),
                  rew_l_b=np.min(train_pot), rew_u_b=np.max(train_pot))
        if type == "train":
            kw.update(eps_ind=self.eps_ind, state_change_penalty=E.state_change_penalty)
        else:
            kw.update(eps_ind=None, state_change_penalty=0.0)
        return kw


def optimum_level_constants(EnvConfig: EnvConfiguration) -> list[dict]:
    """The level-dependent part of ``calculate_optimum`` (rl_opt.py:56-100) for the levels [off, partial, full], in
    the reference's operation order; the hour-dependent part is one multiplication per term (done per hour by the
    host version above or by the CUDA kernel behind ``calculate_optimum_cuda``)."""
    C = EnvConfig
    ms = C.meth_stats_load
    b_s3 = 1 if C.scenario == 3 else 0
    out = []
    for l in (0, 1, 2):
        ch4_volumeflow = ms["Meth_CH4_flow"][l] * C.convert_mol_to_Nm3
        h2_res_volumeflow = ms["Meth_H2_res_flow"][l] * C.convert_mol_to_Nm3
        Q_ch4 = ch4_volumeflow * C.H_u_CH4 * 1000
        Q_h2_res = h2_res_volumeflow * C.H_u_H2 * 1000
        power_chp = Q_ch4 * C.eta_CHP * b_s3
        Q_chp = Q_ch4 * (1 - C.eta_CHP) * b_s3
        Q_steam = ms["Meth_H2O_flow"][l] * (C.dt_water * C.cp_water + C.h_H2O_evap) / 3600
        h2_volumeflow = ms["Meth_H2_flow"][l] * C.convert_mol_to_Nm3
        o2_volumeflow = 1 / 2 * h2_volumeflow * 3600
        Meth_CO2_mass_flow = ms["Meth_CH4_flow"][l] * C.Molar_mass_CO2 / 1000
        eta = electrolyzer_efficiency(h2_volumeflow / C.max_h2_volumeflow, C.min_load_electrolyzer)
        water_elec = ms["Meth_H2_flow"][l] * C.Molar_mass_H2O / 1000 * 3600
        out.append(dict(
            q_gas=Q_ch4 + Q_h2_res, chp_rev=power_chp * C.eeg_el_price, steam_rev=(Q_steam + Q_chp) * C.heat_price,
            o2_rev=o2_volumeflow * C.o2_price, k_eua=Meth_CO2_mass_flow / 1000 * 3600,
            k_heat=ms["Meth_el_heating"][l] / 1000, k_ely=h2_volumeflow * C.H_u_H2 * 1000 / eta,
            water_cost=(ms["Meth_H2O_flow"][l] + water_elec) / C.rho_water * C.water_price,
            stat8=[float(ms[k][l]) for k in ("Meth_State", "Meth_Action", "Meth_Hot_Cold", "Meth_T_cat", "Meth_H2_flow",
                                             "Meth_CH4_flow", "Meth_H2O_flow", "Meth_el_heating")]))
    return out


def calculate_optimum_cuda(el_price_data, gas_price_data, eua_price_data, data_name: str, stats_names,
                           EnvConfig: EnvConfiguration, device="cuda:0", verbose: bool = False) -> dict:
    """``calculate_optimum`` with the per-hour work on the GPU (``ptg_calculate_optimum``): same arguments, same
    24-column dict, bit-identical values.  For sweeps over scenarios / operation points / price series."""
    import ctypes as C_
    import torch
    from . import _abi, _lib
    L = _lib.load()
    dev = torch.device(device)
    el = torch.as_tensor(np.ascontiguousarray(el_price_data, dtype=np.float64)).to(dev)
    gas = torch.as_tensor(np.ascontiguousarray(gas_price_data, dtype=np.float64)).to(dev)
    eua = torch.as_tensor(np.ascontiguousarray(eua_price_data, dtype=np.float64)).to(dev)
    levels = (_abi.PtgOptLevel * 3)()
    for q, lv in enumerate(optimum_level_constants(EnvConfig)):
        for k, v in lv.items():
            if k == "stat8":
                for j, x in enumerate(v):
                    levels[q].stat8[j] = x
            else:
                setattr(levels[q], k, float(v))
    stats = torch.empty((el.numel(), 24), dtype=torch.float64, device=dev)
    stream = C_.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    _lib.check(L.ptg_calculate_optimum(C_.c_void_p(el.data_ptr()), el.numel(), C_.c_void_p(gas.data_ptr()),
                                       C_.c_void_p(eua.data_ptr()), gas.numel(), levels,
                                       C_.c_void_p(stats.data_ptr()), stream))
    host = stats.cpu().numpy()
    stats_dict_opt = {name: host[:, m] for m, name in enumerate(stats_names)}
    if verbose and data_name != "reward_Level":
        print("    > ", data_name, ": Cumulative reward - theoretical optimum T-OPT = ",
              round(stats_dict_opt["Meth_cum_reward_stats"][-EnvConfig.price_ahead], 2))
    return stats_dict_opt
