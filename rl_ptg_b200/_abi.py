"""ctypes mirror of ``include/ptg_b200.h`` and the ``dict_input`` -> ``PtgConfig``/``PtgTables`` marshalling.

``dict_input`` is the flat constructor dict of the reference env (``src/rl_utils.py:337-405``); the reference
splats it into ``self.__dict__`` (``env/ptg_gym_env.py:40``).  Here it is validated and packed into the plain C
structs of the ABI.  Validation errors mirror the reference's ``assert`` sites (``:46, :158, :204``).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from .config import OP_DATASETS

PTG_ABI_VERSION = 3
PTG_N_DATASETS = 17
PTG_N_INFO = 24
PTG_MAX_PRICE_AHEAD = 16
PTG_NCCL_UNIQUE_ID_BYTES = 128

DATASET_NAMES = tuple(k for k, _ in OP_DATASETS)        # index == enum PtgDataset
STATE_NAMES = ("standby", "cooldown", "startup", "partial_load", "full_load")   # index == state/action id

ACT_I64, ACT_I32, ACT_U8, ACT_F32 = 0, 1, 2, 3
NOISE_NUMPY, NOISE_TAPE, NOISE_OFF = 0, 1, 2
SCHED_DUMMY, SCHED_SUBPROC = 0, 1
OBS_KEY_MAJOR, OBS_FLAT = 0, 1

PTG_OK = 0
STATUS_NAMES = {0: "PTG_OK", -1: "PTG_ERR_INVALID_ARGUMENT", -2: "PTG_ERR_CUDA", -3: "PTG_ERR_UNSUPPORTED",
                -4: "PTG_ERR_INVALID_ACTION", -5: "PTG_ERR_DATA_RANGE", -6: "PTG_ERR_NOISE_TAPE", -7: "PTG_ERR_NCCL"}

INFO_KEYS = (  # env/ptg_gym_env.py:253-278, positional order consumed by src/rl_utils.py:544-559
    "step", "el_price_act", "gas_price_act", "eua_price_act", "Meth_State", "Meth_Action", "Meth_Hot_Cold",
    "Meth_T_cat", "Meth_H2_flow", "Meth_CH4_flow", "Meth_H2O_flow", "Meth_el_heating", "ch4_revenues [ct/h]",
    "steam_revenues [ct/h]", "o2_revenues [ct/h]", "eua_revenues [ct/h]", "chp_revenues [ct/h]",
    "elec_costs_heating [ct/h]", "elec_costs_electrolyzer [ct/h]", "water_costs [ct/h]", "reward [ct]",
    "cum_reward", "Pot_Reward", "Part_Full",
)

_I32_FIELDS = (
    "abi_version", "scenario", "raw_modified", "action_type", "train_or_eval", "price_ahead", "sim_step",
    "time_step_op", "eps_sim_steps", "noise_mode", "schedule_mode", "n_eps_loops",
    "time1_start_p_f", "time2_start_f_p", "time_p_f", "time_f_p",
    "time1_p_f_p", "time2_p_f_p", "time23_p_f_p", "time3_p_f_p", "time34_p_f_p", "time4_p_f_p", "time45_p_f_p",
    "time5_p_f_p",
    "time1_f_p_f", "time2_f_p_f", "time23_f_p_f", "time3_f_p_f", "time34_f_p_f", "time4_f_p_f", "time45_f_p_f",
    "time5_f_p_f",
    "i_fully_developed", "j_fully_developed", "obs_layout", "no_auto_reset",
)
_F64_FIELDS = (
    "noise", "eps_len_d", "state_change_penalty", "reward_level",
    "convert_mol_to_Nm3", "H_u_CH4", "H_u_H2", "dt_water", "cp_water", "rho_water", "Molar_mass_CO2",
    "Molar_mass_H2O", "h_H2O_evap", "eeg_el_price", "heat_price", "o2_price", "water_price",
    "min_load_electrolyzer", "max_h2_volumeflow", "eta_CHP",
    "t_cat_standby", "t_cat_startup_cold", "t_cat_startup_hot",
    "el_l_b", "el_u_b", "gas_l_b", "gas_u_b", "eua_l_b", "eua_u_b", "T_l_b", "T_u_b", "h2_l_b", "h2_u_b",
    "ch4_l_b", "ch4_u_b", "h2_res_l_b", "h2_res_u_b", "h2o_l_b", "h2o_u_b", "heat_l_b", "heat_u_b",
    "rew_l_b", "rew_u_b",
)
_TIME_KEYS = tuple(f for f in _I32_FIELDS if f.startswith("time") and f != "time_step_op") + (
    "i_fully_developed", "j_fully_developed")
_DIRECT_F64 = tuple(f for f in _F64_FIELDS if f not in ("reward_level",))


class PtgConfig(C.Structure):
    _fields_ = [(f, C.c_int32) for f in _I32_FIELDS] + [(f, C.c_double) for f in _F64_FIELDS]


class PtgTables(C.Structure):
    _fields_ = [
        ("op", C.c_void_p * PTG_N_DATASETS),
        ("op_rows", C.c_int64 * PTG_N_DATASETS),
        ("e_r_b", C.c_void_p), ("n_hours", C.c_int64),
        ("g_e", C.c_void_p), ("n_days", C.c_int64),
        ("eps_ind", C.c_void_p), ("n_eps_ind", C.c_int64),
    ]


class PtgOptLevel(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("q_gas", "chp_rev", "steam_rev", "o2_rev", "k_eua", "k_heat", "k_ely",
                                          "water_cost")] + [("stat8", C.c_double * 8)]


class PtgIO(C.Structure):
    _fields_ = [("obs", C.c_void_p), ("reward", C.c_void_p), ("done", C.c_void_p), ("terminal_obs", C.c_void_p),
                ("info", C.c_void_p), ("episode_return", C.c_void_p), ("episode_length", C.c_void_p),
                ("windows_changed", C.c_void_p), ("status_u8", C.c_void_p)]


class PtgObsKey(C.Structure):
    _fields_ = [("name", C.c_char * 24), ("dim", C.c_int32), ("is_int32", C.c_int32), ("offset", C.c_int64)]


class PtgEpisodeStats(C.Structure):
    _fields_ = [(f, C.c_double) for f in ("count", "sum_return", "sum_return_sq", "sum_length", "min_return",
                                          "max_return", "total_steps", "_reserved")]


_STATE_I32 = ("meth_state", "i", "j", "k", "hot_cold", "standby_ds", "startup_ds", "partial_ds", "full_ds",
              "current_action", "act_ep_h", "act_ep_d", "episode_count")


class PtgStateSoA(C.Structure):
    _fields_ = [(f, C.c_void_p) for f in _STATE_I32] + [("draws", C.c_void_p), ("t_cat", C.c_void_p),
                                                        ("cum_reward", C.c_void_p), ("rng", C.c_void_p),
                                                        ("state_changes", C.c_void_p)]


STATE_FIELDS = tuple((f, np.int32) for f in _STATE_I32) + (("draws", np.int64), ("t_cat", np.float64),
                                                           ("cum_reward", np.float64), ("rng", np.uint64),
                                                           ("state_changes", np.uint32))
STATE_WIDTH = {"rng": 4}        # fields with more than one value per env: rng = PCG64 {state_hi, state_lo, inc_hi, inc_lo}


def state_shape(name: str, n_envs: int) -> tuple:
    return (n_envs, STATE_WIDTH[name]) if name in STATE_WIDTH else (n_envs,)


def alloc_state(n_envs: int):
    """Host arrays + the struct pointing at them."""
    arrays = {name: np.zeros(state_shape(name, n_envs), dtype=dt) for name, dt in STATE_FIELDS}
    s = PtgStateSoA()
    for name, _ in STATE_FIELDS:
        setattr(s, name, arrays[name].ctypes.data)
    return s, arrays


def _as_int(name: str, v) -> int:
    iv = int(v)
    if iv != v:
        raise ValueError(f"config knob {name} = {v!r} must be integral")
    return iv


def config_from_kwargs(dict_input: dict, train_or_eval: str = "train", noise_mode: int = NOISE_NUMPY,
                       schedule_mode: int | None = None, obs_layout: int = 0, auto_reset: bool = True) -> PtgConfig:
    """Validate the reference constructor dict and pack its scalars."""
    d = dict_input
    if train_or_eval not in ("train", "eval"):
        raise ValueError('train_or_eval must be either "train" or "eval".')        # ptg_gym_env.py:46
    if d["action_type"] not in ("discrete", "continuous"):
        raise ValueError(f"invalid action type ({d['action_type']}) - must match ['discrete', 'continuous']!")
    if d["raw_modified"] not in ("raw", "mod"):
        raise ValueError(f"state design raw_modified {d['raw_modified']} must match 'raw' or 'mod'!")
    for sid, name in enumerate(STATE_NAMES):
        # the reference indexes list(M_state.keys()) with the state id (:109,:359) -> ids are positional
        if int(d[f"ptg_{name}"]) != sid:
            raise ValueError(f"ptg_{name} must be {sid} (the reference's state list is positional)")
    cfg = PtgConfig()
    cfg.abi_version = PTG_ABI_VERSION
    cfg.scenario = _as_int("scenario", d["scenario"])
    cfg.raw_modified = 1 if d["raw_modified"] == "mod" else 0
    cfg.action_type = 1 if d["action_type"] == "continuous" else 0
    cfg.train_or_eval = 1 if train_or_eval == "eval" else 0
    cfg.price_ahead = _as_int("price_ahead", d["price_ahead"])
    cfg.sim_step = _as_int("sim_step", d["sim_step"])
    cfg.time_step_op = _as_int("time_step_op", d["time_step_op"])
    cfg.eps_sim_steps = _as_int("eps_sim_steps", d["eps_sim_steps"])
    cfg.noise_mode = noise_mode
    if schedule_mode is None:
        schedule_mode = SCHED_SUBPROC if d.get("parallel") == "Multiprocessing" else SCHED_DUMMY
    cfg.schedule_mode = schedule_mode
    cfg.obs_layout = int(obs_layout)
    cfg.no_auto_reset = 0 if auto_reset else 1
    cfg.n_eps_loops = max(1, int(d.get("n_eps_loops", 1) or 1))
    for k in _TIME_KEYS:
        setattr(cfg, k, _as_int(k, d[k]))
    for k in _DIRECT_F64:
        setattr(cfg, k, float(d[k]))
    cfg.reward_level = float(np.asarray(d["reward_level"], dtype=np.float64).reshape(-1)[0])   # r_0, :125
    return cfg


def tables_from_kwargs(dict_input: dict, price_ahead: int):
    """Contiguous fp64 views of the dict's arrays.  Returns ``(PtgTables, keepalive)``."""
    d = dict_input
    keep = []
    t = PtgTables()
    for idx, name in enumerate(DATASET_NAMES):
        a = np.ascontiguousarray(d[name], dtype=np.float64)
        if a.ndim != 2 or a.shape[1] != 7:
            raise ValueError(f"operation table {name} must have shape [rows, 7], got {a.shape}")
        keep.append(a)
        t.op[idx] = a.ctypes.data
        t.op_rows[idx] = a.shape[0]
    e_r_b = np.ascontiguousarray(d["e_r_b"], dtype=np.float64)
    if e_r_b.ndim != 3 or e_r_b.shape[0] != 3 or e_r_b.shape[1] != price_ahead:
        raise ValueError(f"e_r_b must have shape [3, price_ahead={price_ahead}, hours], got {e_r_b.shape}")
    g_e = np.ascontiguousarray(d["g_e"], dtype=np.float64)
    if g_e.ndim != 3 or g_e.shape[:2] != (2, 2):
        raise ValueError(f"g_e must have shape [2, 2, days], got {g_e.shape}")
    keep += [e_r_b, g_e]
    t.e_r_b, t.n_hours = e_r_b.ctypes.data, e_r_b.shape[2]
    t.g_e, t.n_days = g_e.ctypes.data, g_e.shape[2]
    eps_ind = d.get("eps_ind")
    if isinstance(eps_ind, np.ndarray):      # training env (ptg_gym_env.py:59)
        e = np.ascontiguousarray(eps_ind, dtype=np.int64)
        keep.append(e)
        t.eps_ind, t.n_eps_ind = e.ctypes.data, e.shape[0]
    else:
        t.eps_ind, t.n_eps_ind = None, 0
    return t, keep


def obs_keys(raw_modified: str, price_ahead: int):
    """(name, dim, is_int) in the key order of the reference obs dict (ptg_gym_env.py:222-249)."""
    if raw_modified == "raw":
        head = [("Elec_Price", price_ahead, False), ("Gas_Price", 2, False), ("EUA_Price", 2, False)]
    else:
        head = [("Pot_Reward", price_ahead, False), ("Part_Full", price_ahead, False)]
    tail = [("METH_STATUS", 1, True)] + [(k, 1, False) for k in (
        "T_CAT", "H2_in_MolarFlow", "CH4_syn_MolarFlow", "H2_res_MolarFlow", "H2O_DE_MassFlow", "Elec_Heating",
        "Temp_hour_enc_sin", "Temp_hour_enc_cos")]
    return head + tail
