"""Market + process data loading for the batched PtG environment.

Mirrors ``load_data`` of the reference (``src/rl_utils.py:21-144``): same CSV format (``;``-delimited, columns
by name), same unit conversions (EUR/MWh -> ct/kWh for electricity and gas, ``:35,:37``), same BS2/BS3 price
overrides (``:119-126``), same reward-level entries (``:129-130``) and the same divisibility checks
(``:133-142``).  Additionally provides ``synthetic_data`` -- seeded market/process data of the repo's shapes
 for boxes where the reference data set is not mounted -- and an ``.npz`` round trip so real
data can travel as a compact fixture.
"""
from __future__ import annotations

import os
import warnings

import numpy as np

from .config import OP_DATASETS, EnvConfiguration, TrainConfiguration

OP_COLUMNS = ("Time [s]", "T_cat [gradC]", "n_h2 [mol/s]", "n_ch4 [mol/s]", "n_h2_res [mol/s]", "m_DE [kg/h]",
              "Pel [W]")
MARKET_COLUMNS = {"el": ("Day-Ahead-price [Euro/MWh]", 10.0), "gas": ("THE_DA_Gas [Euro/MWh]", 10.0),
                  "eua": ("EUA_CO2 [Euro/t]", None)}
SPLITS = ("train", "val", "test")
MIN_TRAIN_LEN_D = 6  # days kept free at the end of the training set for the day-ahead window (rl_utils.py:133)


def _read_csv_columns(file_path: str, columns) -> list[np.ndarray]:
    import pandas as pd
    df = pd.read_csv(file_path, delimiter=";", decimal=".")
    return [df[c].values.astype(float) for c in columns]


def import_market_data(csvfile: str, type: str, path: str) -> np.ndarray:
    """One day-ahead price series (reference ``import_market_data``, src/rl_utils.py:21-43)."""
    if type not in MARKET_COLUMNS:
        raise ValueError("Invalid market data type. Must be one of ['el', 'gas', 'eua']!")
    col, div = MARKET_COLUMNS[type]
    (arr,) = _read_csv_columns(path + "/" + csvfile, [col])
    return arr / div if div is not None else arr


def import_data(csvfile: str, path: str) -> np.ndarray:
    """One methanation time-series table ``[rows, 7]`` (reference ``import_data``, src/rl_utils.py:46-67)."""
    return np.stack(_read_csv_columns(path + "/" + csvfile, OP_COLUMNS), axis=1)


def _finish_price_dict(dict_price_data: dict, EnvConfig: EnvConfiguration) -> None:
    """Scenario overrides, reward-level entries and training-set checks (src/rl_utils.py:119-142)."""
    for split in SPLITS:
        el_h = len(dict_price_data[f"el_price_{split}"])
        sizes = {el_h // 24, len(dict_price_data[f"gas_price_{split}"]), len(dict_price_data[f"eua_price_{split}"])}
        if len(sizes) > 1:
            warnings.warn(f"Market data size does not match for {split}: electricity ({el_h}h), gas/EUA days "
                          f"{sorted(sizes)} -> Check size!", UserWarning)
    if EnvConfig.scenario in (2, 3):
        gas_price = EnvConfig.ch4_price_fix if EnvConfig.scenario == 2 else 0
        for split in SPLITS:
            dict_price_data[f"gas_price_{split}"] = np.full(len(dict_price_data[f"gas_price_{split}"]), gas_price)
        if EnvConfig.scenario == 3:
            for split in SPLITS:
                dict_price_data[f"eua_price_{split}"] = np.zeros(len(dict_price_data[f"eua_price_{split}"]))
    for key in ("el_price", "gas_price", "eua_price"):
        dict_price_data[f"{key}_reward_level"] = EnvConfig.r_0_values[key]

    EnvConfig.train_len_d = len(dict_price_data["gas_price_train"]) - MIN_TRAIN_LEN_D
    if EnvConfig.train_len_d <= 0:
        raise ValueError(f"The training set size must be greater than {MIN_TRAIN_LEN_D} days")
    if EnvConfig.train_len_d % EnvConfig.eps_len_d != 0:
        divisors = [i for i in range(1, EnvConfig.train_len_d + 1) if EnvConfig.train_len_d % i == 0]
        raise ValueError(f"The training set size {EnvConfig.train_len_d} must be divisible by the episode length "
                         f"eps_len_d : {EnvConfig.eps_len_d}; possible divisors are: {divisors}")


def load_data(EnvConfig: EnvConfiguration, TrainConfig: TrainConfiguration):
    """CSV files -> ``(dict_price_data, dict_op_data)``; same keys as the reference (src/rl_utils.py:70-144)."""
    path = TrainConfig.path
    if path is None:
        raise ValueError("TrainConfig.path (the data root holding data/...) is not set")
    dict_price_data = {
        f"{name}_price_{split}": import_market_data(getattr(EnvConfig, f"datafile_path_{split}_{name}"), name, path)
        for name in ("el", "gas", "eua") for split in SPLITS
    }
    dict_op_data = {key: import_data(getattr(EnvConfig, f"datafile_path{num}"), path) for key, num in OP_DATASETS}
    _finish_price_dict(dict_price_data, EnvConfig)
    return dict_price_data, dict_op_data


# ------------------------------------------------------------------------------------------------------------
# compact .npz round trip (raw series, before scenario overrides)
# ------------------------------------------------------------------------------------------------------------
def save_raw_npz(file_path: str, data_root: str, operations=("OP1", "OP2")) -> None:
    """Pack the raw CSV content (no scenario overrides) into one compressed ``.npz``."""
    out = {}
    cfg = EnvConfiguration()
    for name in ("el", "gas", "eua"):
        for split in SPLITS:
            out[f"market/{name}_{split}"] = import_market_data(getattr(cfg, f"datafile_path_{split}_{name}"), name,
                                                               data_root)
    for op in operations:
        c = EnvConfiguration(operation=op)
        for key, num in OP_DATASETS:
            out[f"{op}/{key}"] = import_data(getattr(c, f"datafile_path{num}"), data_root)
    np.savez_compressed(file_path, **out)


def load_data_npz(file_path: str, EnvConfig: EnvConfiguration):
    """Inverse of ``save_raw_npz`` + the scenario handling of ``load_data``."""
    with np.load(file_path) as z:
        dict_price_data = {f"{name}_price_{split}": np.array(z[f"market/{name}_{split}"], dtype=np.float64)
                           for name in ("el", "gas", "eua") for split in SPLITS}
        dict_op_data = {key: np.array(z[f"{EnvConfig.operation}/{key}"], dtype=np.float64) for key, _ in OP_DATASETS}
    _finish_price_dict(dict_price_data, EnvConfig)
    return dict_price_data, dict_op_data


# ------------------------------------------------------------------------------------------------------------
# synthetic data of the repo's shapes
# ------------------------------------------------------------------------------------------------------------
_OP_ROWS = {
    "OP2": dict(cooldown=45001, standby_down=13829, standby_up=15546, startup_cold=3422, startup_hot=2048,
                op1_start_p=4987, op2_start_f=2223, op3_p_f=3015, op4_p_f_p_5=2780, op5_p_f_p_10=3021,
                op6_p_f_p_15=3261, op7_p_f_p_22=3481, op8_f_p=2820, op9_f_p_f_5=2947, op10_f_p_f_10=3297,
                op11_f_p_f_15=3466, op12_f_p_f_20=3616),
    "OP1": dict(cooldown=45001, standby_down=13829, standby_up=15546, startup_cold=3422, startup_hot=2048,
                op1_start_p=6678, op2_start_f=4987, op3_p_f=2438, op4_p_f_p_5=3100, op5_p_f_p_10=4041,
                op6_p_f_p_15=4041, op7_p_f_p_22=4785, op8_f_p=6452, op9_f_p_f_5=2976, op10_f_p_f_10=3215,
                op11_f_p_f_15=3316, op12_f_p_f_20=3907),
}


def _ramp(n, a, b, rng, wiggle):
    """Smooth monotone-ish ramp a -> b with a small band-limited wiggle."""
    x = np.linspace(0.0, 1.0, n)
    base = a + (b - a) * (1.0 - np.exp(-4.0 * x)) / (1.0 - np.exp(-4.0))
    w = np.cumsum(rng.normal(0.0, 1.0, n))
    w -= np.linspace(w[0], w[-1], n)
    return base + wiggle * w / (np.abs(w).max() + 1e-12)


def synthetic_op_data(operation: str = "OP2", seed: int = 0) -> dict:
    """17 methanation tables ``[rows, 7]`` with the row counts and value ranges of the reference data."""
    rng = np.random.default_rng(seed)
    rows = _OP_ROWS[operation]
    part, full = (0.0198, 0.0485) if operation == "OP2" else (0.00701, 0.0198)
    out = {}

    def table(n, T, h2_level, heat):
        t = np.arange(n, dtype=np.float64) * 2.0
        T = np.round(np.clip(T, 0.0, 600.0), 1)
        h2 = np.round(np.clip(h2_level, 0.0, None), 5)
        conv = np.clip((T - 200.0) / 250.0, 0.0, 1.0) * (h2 > 0)
        ch4 = np.round(h2 * 0.235 * conv, 5)
        h2res = np.round(h2 * 0.035 * conv * (0.3 + 0.7 * rng.random(n)), 6)
        h2o = np.round(ch4 * 107.0, 4)
        pel = np.round(np.clip(heat, 0.0, 1750.0), 0)
        return np.stack([t, T, h2, ch4, h2res, h2o, pel], axis=1)

    n = rows["cooldown"]
    out["cooldown"] = table(n, _ramp(n, 598.7, 11.0, rng, 1.5), np.zeros(n), np.zeros(n))
    n = rows["standby_down"]
    out["standby_down"] = table(n, _ramp(n, 560.0, 190.0, rng, 2.0), np.zeros(n), 150.0 + 60.0 * rng.random(n))
    n = rows["standby_up"]
    out["standby_up"] = table(n, _ramp(n, 15.0, 188.4, rng, 1.0), np.zeros(n), 900.0 + 300.0 * rng.random(n))
    n = rows["startup_cold"]
    out["startup_cold"] = table(n, _ramp(n, 12.0, 400.0, rng, 3.0), part * np.clip(np.linspace(-1.5, 1, n), 0, 1),
                                1400.0 + 300.0 * rng.random(n))
    n = rows["startup_hot"]
    out["startup_hot"] = table(n, _ramp(n, 170.0, 405.0, rng, 3.0), part * np.clip(np.linspace(-0.5, 1, n), 0, 1),
                               1200.0 + 300.0 * rng.random(n))
    for key in rows:
        if key in out:
            continue
        n = rows[key]
        to_full = key in ("op2_start_f", "op3_p_f", "op9_f_p_f_5", "op10_f_p_f_10", "op11_f_p_f_15",
                          "op12_f_p_f_20")
        lo, hi = (part, full) if to_full else (full, part)
        frac = np.clip(np.linspace(0.0, 4.0, n), 0.0, 1.0)
        h2 = lo + (hi - lo) * frac
        T0, T1 = (420.0, 505.0) if to_full else (500.0, 425.0)
        out[key] = table(n, _ramp(n, T0, T1, rng, 4.0), h2, 330.0 + 60.0 * rng.random(n))
    return {key: out[key] for key, _ in OP_DATASETS}


def synthetic_price_data(EnvConfig: EnvConfiguration, seed: int = 0, train_days: int = 1523, val_days: int = 61,
                         test_days: int = 61) -> dict:
    """AR(1) series with the moments of the reference market data (el: mean 11.8, sd 10.8 ct/kWh, ...)."""
    rng = np.random.default_rng(seed + 1)
    out = {}

    def ar1(n, mean, sd, rho, lo, hi):
        e = rng.normal(0.0, 1.0, n)
        x = np.empty(n)
        x[0] = e[0]
        c = np.sqrt(1.0 - rho * rho)
        for i in range(1, n):
            x[i] = rho * x[i - 1] + c * e[i]
        return np.clip(mean + sd * x, lo, hi)

    for split, days in (("train", train_days), ("val", val_days), ("test", test_days)):
        hours = days * 24 + (0 if split != "val" else 1)   # the reference val set has 1465 = 61*24+1 hours
        out[f"el_price_{split}"] = np.round(ar1(hours, 118.0, 108.0, 0.97, -500.0, 871.0), 2) / 10
        out[f"gas_price_{split}"] = np.round(ar1(days, 60.0, 30.0, 0.995, 22.0, 128.0), 3) / 10
        out[f"eua_price_{split}"] = np.round(ar1(days, 70.0, 20.0, 0.995, 12.9, 121.0), 3)
    _finish_price_dict(out, EnvConfig)
    return out


def synthetic_data(EnvConfig: EnvConfiguration, seed: int = 0):
    """``(dict_price_data, dict_op_data)`` without touching the file system."""
    return synthetic_price_data(EnvConfig, seed), synthetic_op_data(EnvConfig.operation, seed)
