"""B200-native batched simulator for the SSE hot path of the quantum-cartpole environments.

Public surface:
  BatchedSim           one libqcart handle: set_state / step / get_moments on B trajectories (sim.py)
  QuantumCartpoleEnv   reset/step environment API with the reference's observation, reward and termination rules (env.py)
  simulation           drop-in mirror of the reference's compiled `simulation` module (simulation.py)
  controllers          analytic controllers (LQG / damping / semiclassical) + 21-level quantisation from the GPU moment block (controllers.py)
  rollout              device-resident policy (direct_DQN), epsilon-greedy, experience rows, measurement record (rollout.py)
  configs              the four task presets + the grid-size sweep (configs.py)
The CUDA library (csrc/, built to libqcart.so) is required: there is no CPU fallback.
"""
from . import configs  # noqa: F401
from ._lib import QcartError, LIB_PATH  # noqa: F401
from .sim import BatchedSim, philox_normals, measure_peaks, make_config  # noqa: F401
from .env import QuantumCartpoleEnv  # noqa: F401
from . import simulation  # noqa: F401
from . import controllers  # noqa: F401
from . import rollout  # noqa: F401

__all__ = ["BatchedSim", "QuantumCartpoleEnv", "simulation", "configs", "philox_normals", "measure_peaks", "QcartError"]
