"""Shared test helpers: golden fixtures -> env kwargs via the PRODUCT's own preprocessing."""
from __future__ import annotations

import functools
import glob
import json
import os

import numpy as np

import rl_ptg_b200 as ptg

ROOT_DIR = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
REF_DATA = os.path.join(GOLDEN_DIR, "ref_data.npz")
GOLDEN_CASES = sorted(os.path.basename(p)[len("golden_"):-len(".npz")]
                      for p in glob.glob(os.path.join(GOLDEN_DIR, "golden_*.npz")))


@functools.lru_cache(maxsize=None)
def load_golden(case: str) -> dict:
    with np.load(os.path.join(GOLDEN_DIR, f"golden_{case}.npz"), allow_pickle=False) as z:
        g = {k: z[k] for k in z.files}
    g["meta"] = json.loads(str(g["meta"]))
    return g


@functools.lru_cache(maxsize=None)
def _preprocess(overrides_json: str, action_type: str, seed_train: int):
    overrides = json.loads(overrides_json)
    E = ptg.EnvConfiguration(**overrides)
    T = ptg.TrainConfiguration(seed_train=seed_train)
    A = ptg.AgentConfiguration(action_type=action_type)
    price, op = ptg.load_data_npz(REF_DATA, E)
    return ptg.Preprocessing(price, op, A, E, T)


def real_kwargs(overrides: dict | None = None, split: str = "train", action_type: str = "discrete",
                seed_train: int = 3654) -> dict:
    """Env kwargs on the REAL reference data (tests/golden/ref_data.npz) built by the product's preprocessing."""
    pp = _preprocess(json.dumps(overrides or {}, sort_keys=True), action_type, seed_train)
    return pp.dict_env_kwargs(split)


def golden_kwargs(case: str) -> dict:
    m = load_golden(case)["meta"]
    return real_kwargs(m["overrides"], m["split"], m["action_type"], m["seed_train"])


@functools.lru_cache(maxsize=None)
def _preprocess_synth(overrides_json: str, action_type: str, seed: int):
    overrides = json.loads(overrides_json)
    E = ptg.EnvConfiguration(**overrides)
    T = ptg.TrainConfiguration()
    A = ptg.AgentConfiguration(action_type=action_type)
    price, op = ptg.synthetic_data(E, seed)
    return ptg.Preprocessing(price, op, A, E, T)


def synthetic_kwargs(overrides: dict | None = None, split: str = "train", action_type: str = "discrete",
                     seed: int = 0) -> dict:
    return _preprocess_synth(json.dumps(overrides or {}, sort_keys=True), action_type, seed).dict_env_kwargs(split)


def obs_rel_err(a: np.ndarray, b: np.ndarray) -> float:
    """max |a-b| / max(|b|, tiny) with exact zeros required to match exactly."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    denom = np.maximum(np.abs(b), 1e-30)
    return float(np.max(np.abs(a - b) / denom)) if a.size else 0.0
