"""Observation/action spaces of the env.  Uses gymnasium's classes when gymnasium is importable, else minimal
stand-ins with the same constructor arguments and attributes (``shape``, ``dtype``, ``low``, ``high``, ``n``,
``spaces``, ``sample``, ``contains``) -- enough for SB3-style consumers that only introspect the spaces."""
from __future__ import annotations

import numpy as np

try:  # pragma: no cover - not installed in the build container
    from gymnasium.spaces import Box, Dict, Discrete  # type: ignore  # noqa: F401
    HAVE_GYMNASIUM = True
except ModuleNotFoundError:
    HAVE_GYMNASIUM = False

    class _Space:
        shape = None
        dtype = None

        def seed(self, seed=None):
            self._rng = np.random.default_rng(seed)

        @property
        def np_random(self):
            if not hasattr(self, "_rng"):
                self._rng = np.random.default_rng()
            return self._rng

    class Box(_Space):
        def __init__(self, low, high, shape=None, dtype=np.float32):
            self.dtype = np.dtype(dtype)
            if shape is None:
                shape = np.shape(low)
            self.shape = tuple(shape)
            self.low = np.broadcast_to(np.asarray(low, dtype=self.dtype), self.shape).copy()
            self.high = np.broadcast_to(np.asarray(high, dtype=self.dtype), self.shape).copy()

        def sample(self):
            return self.np_random.uniform(self.low, self.high).astype(self.dtype)

        def contains(self, x):
            x = np.asarray(x)
            return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

        def __repr__(self):
            return f"Box({self.low.min()}, {self.high.max()}, {self.shape}, {self.dtype})"

    class Discrete(_Space):
        def __init__(self, n, start=0):
            self.n, self.start = int(n), int(start)
            self.shape, self.dtype = (), np.dtype(np.int64)

        def sample(self):
            return int(self.np_random.integers(self.start, self.start + self.n))

        def contains(self, x):
            return self.start <= int(x) < self.start + self.n

        def __repr__(self):
            return f"Discrete({self.n})"

    class Dict(_Space):
        def __init__(self, spaces=None):
            self.spaces = dict(spaces or {})

        def keys(self):
            return self.spaces.keys()

        def items(self):
            return self.spaces.items()

        def __getitem__(self, k):
            return self.spaces[k]

        def __iter__(self):
            return iter(self.spaces)

        def __len__(self):
            return len(self.spaces)

        def sample(self):
            return {k: s.sample() for k, s in self.spaces.items()}

        def contains(self, x):
            return set(x.keys()) == set(self.spaces.keys()) and all(self.spaces[k].contains(x[k]) for k in x)

        def __repr__(self):
            return "Dict(" + ", ".join(f"{k!r}: {s!r}" for k, s in self.spaces.items()) + ")"


def observation_space(raw_modified: str, price_ahead: int) -> "Dict":
    """env/ptg_gym_env.py:160-204 -- same keys, shapes and bounds; all Boxes float64 like the reference."""
    b_norm, b_enc = (0, 1), (-1, 1)
    pa = price_ahead
    box1 = lambda b: Box(low=b[0], high=b[1], shape=(1,), dtype=np.float64)      # noqa: E731
    vec = lambda b, n: Box(low=b[0] * np.ones((n,)), high=b[1] * np.ones((n,)), dtype=np.float64)   # noqa: E731
    if raw_modified == "raw":
        head = {"Elec_Price": vec(b_norm, pa), "Gas_Price": vec(b_norm, 2), "EUA_Price": vec(b_norm, 2)}
    elif raw_modified == "mod":
        head = {"Pot_Reward": vec(b_norm, pa), "Part_Full": vec(b_enc, pa)}
    else:
        raise ValueError(f"state design raw_modified {raw_modified} must match 'raw' or 'mod'!")
    tail = {"METH_STATUS": Discrete(6)}
    for k in ("T_CAT", "H2_in_MolarFlow", "CH4_syn_MolarFlow", "H2_res_MolarFlow", "H2O_DE_MassFlow", "Elec_Heating"):
        tail[k] = box1(b_norm)
    tail["Temp_hour_enc_sin"] = box1(b_enc)
    tail["Temp_hour_enc_cos"] = box1(b_enc)
    return Dict({**head, **tail})


def action_space(action_type: str):
    """env/ptg_gym_env.py:140-158."""
    if action_type == "discrete":
        return Discrete(5)
    if action_type == "continuous":
        return Box(low=-1, high=1, shape=(1,), dtype=np.float32)
    raise ValueError(f"invalid action type ({action_type}) - must match ['discrete', 'continuous']!")
