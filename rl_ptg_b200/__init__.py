"""ptg-b200: a B200-native batched implementation of RL_PtG's PTGEnv.step()/reset() hot path."""
from .config import AgentConfiguration, EnvConfiguration, TrainConfiguration  # noqa: F401
from .data import load_data, load_data_npz, synthetic_data  # noqa: F401
from .preprocessing import Preprocessing, calculate_optimum  # noqa: F401

__version__ = "0.1.0"
