"""Reference harness: run the UNMODIFIED SimMarkt/RL_PtG code in-process (TEST INFRASTRUCTURE ONLY).

This file is part of ``oracle/`` -- it is a checker, never shipped on the product path.  It only works in a
container where the reference checkout is mounted (``/root/reference``); it is used by
``tests/golden/gen_golden.py`` to produce the committed golden vectors and is never imported by ``-m gpu``
tests, ``smoke()`` or ``bench.py``.

``gymnasium``, ``stable_baselines3`` and ``matplotlib`` are not installed here, and the reference imports
them at module top (env/ptg_gym_env.py:4-5, src/rl_utils.py:11,15-17).  We install small ``sys.modules``
stubs (the recipe of SURVEY.md Appendix C) -- the stub ``gymnasium.Env`` reproduces the only behaviour the
env relies on: ``np_random`` = ``Generator(PCG64(SeedSequence(seed)))`` re-created by ``reset(seed=...)``.
"""
from __future__ import annotations

import contextlib
import os
import shutil
import sys
import tempfile
import types

import numpy as np

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STAGED_ROOT = os.path.join(_REPO, "baseline", "_ref")      # tools/stage_reference.sh: travels to the GPU box


def _default_root() -> str:
    """The mounted checkout in the build container, else the staged copy (the only one a GPU box has)."""
    for cand in (os.environ.get("PTG_REFERENCE_ROOT"), "/root/reference", STAGED_ROOT):
        if cand and os.path.isfile(os.path.join(cand, "env", "ptg_gym_env.py")):
            return cand
    return "/root/reference"


REF_ROOT = _default_root()


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "env", "ptg_gym_env.py"))


def _install_stubs() -> None:
    sys.dont_write_bytecode = True  # the reference mount is read-only
    try:
        import gymnasium  # noqa: F401  (prefer the real package when present)
    except ModuleNotFoundError:
        gym = types.ModuleType("gymnasium")

        class Env:
            _np_random = None

            @property
            def np_random(self):
                if self._np_random is None:
                    self._np_random = np.random.Generator(np.random.PCG64(np.random.SeedSequence()))
                return self._np_random

            def reset(self, seed=None, options=None):
                if seed is not None:
                    self._np_random = np.random.Generator(np.random.PCG64(np.random.SeedSequence(seed)))

        class _Space:
            def __init__(self, *a, **k):
                self.args, self.kwargs = a, k
                self.shape = k.get("shape")
                self.dtype = k.get("dtype")

        class Box(_Space):
            pass

        class Discrete(_Space):
            def __init__(self, n, *a, **k):
                super().__init__(n, *a, **k)
                self.n = n

        class Dict(_Space):
            def __init__(self, spaces=None, **k):
                super().__init__(spaces, **k)
                self.spaces = spaces

        spaces = types.ModuleType("gymnasium.spaces")
        spaces.Box, spaces.Discrete, spaces.Dict = Box, Discrete, Dict
        gym.Env, gym.spaces = Env, spaces
        gym.register = lambda *a, **k: None
        sys.modules["gymnasium"] = gym
        sys.modules["gymnasium.spaces"] = spaces
    for name in ("matplotlib", "matplotlib.pyplot"):
        try:
            __import__(name)
        except ModuleNotFoundError:
            sys.modules[name] = types.ModuleType(name)
    try:
        import stable_baselines3  # noqa: F401
    except ModuleNotFoundError:
        for name, attrs in (
            ("stable_baselines3", ()),
            ("stable_baselines3.common", ()),
            ("stable_baselines3.common.vec_env", ("VecNormalize", "DummyVecEnv", "SubprocVecEnv")),
            ("stable_baselines3.common.env_util", ("make_vec_env",)),
            ("stable_baselines3.common.callbacks", ("EvalCallback",)),
        ):
            m = types.ModuleType(name)
            for a in attrs:
                setattr(m, a, type(a, (), {}))
            sys.modules[name] = m


@contextlib.contextmanager
def _chdir(path):
    old = os.getcwd()
    os.chdir(path)
    try:
        yield
    finally:
        os.chdir(old)


class ReferenceSession:
    """Imports the reference and builds its env kwargs for one config variant.

    ``overrides`` are textual ``key : value`` replacements in a temp copy of ``config/config_env.yaml`` (the
    reference re-reads that file from the cwd in ``calculate_optimum``, src/rl_opt.py:37).
    """

    def __init__(self, overrides: dict | None = None, action_type: str = "discrete",
                 seed_train: int = 3654, seed_test: int = 605, train_steps: int | None = None):
        assert reference_available(), f"reference not mounted at {REF_ROOT}"
        _install_stubs()
        if REF_ROOT not in sys.path:
            sys.path.insert(0, REF_ROOT)
        self.tmp = tempfile.mkdtemp(prefix="ptg_ref_cfg_")
        os.makedirs(os.path.join(self.tmp, "config"))
        for f in ("config_agent.yaml", "config_train.yaml"):
            shutil.copy(os.path.join(REF_ROOT, "config", f), os.path.join(self.tmp, "config", f))
        with open(os.path.join(REF_ROOT, "config", "config_env.yaml")) as fh:
            lines = fh.read().split("\n")
        overrides = dict(overrides or {})
        out = []
        for ln in lines:
            key = ln.split(":")[0].strip() if ":" in ln and not ln.startswith((" ", "#")) else None
            if key in overrides:
                out.append(f"{key} : {overrides.pop(key)}")
            else:
                out.append(ln)
        assert not overrides, f"unknown config_env keys: {list(overrides)}"
        with open(os.path.join(self.tmp, "config", "config_env.yaml"), "w") as fh:
            fh.write("\n".join(out))

        with _chdir(self.tmp):
            from src.rl_config_agent import AgentConfiguration
            from src.rl_config_env import EnvConfiguration
            from src.rl_config_train import TrainConfiguration
            import src.rl_utils as ru
            import env.ptg_gym_env as pg
            self.pg = pg
            A, E, T = AgentConfiguration(), EnvConfiguration(), TrainConfiguration()
            T.path = REF_ROOT
            T.seed_train, T.seed_test = seed_train, seed_test
            if train_steps is not None:
                T.train_steps = train_steps
            A.rl_alg_hyp["action_type"] = action_type
            import io
            with contextlib.redirect_stdout(io.StringIO()) as buf:
                self.dict_price_data, self.dict_op_data = ru.load_data(E, T)
                self.P = ru.Preprocessing(self.dict_price_data, self.dict_op_data, A, E, T)
            self.stdout = buf.getvalue()
            self.A, self.E, self.T = A, E, T

    def kwargs(self, split: str = "train") -> dict:
        return self.P.dict_env_kwargs(split)

    def make_env(self, split: str = "train", train_or_eval: str = "train", kw: dict | None = None):
        """Fresh reference PTGEnv with the module-global episode counter reset to 0 (ptg_gym_env.py:9)."""
        self.pg.ep_index = 0
        return self.pg.PTGEnv(kw if kw is not None else self.kwargs(split), train_or_eval)

    def close(self):
        shutil.rmtree(self.tmp, ignore_errors=True)
