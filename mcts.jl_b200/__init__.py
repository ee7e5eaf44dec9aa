"""mcts.jl_b200 -- B200-native search loop behind MCTS.jl's MCTSSolver/DPWSolver API.

The directory name contains a dot, so import it through the top-level shim: `import mcts_jl_b200`.

Layout:
  csrc/        hand-written sm_100a CUDA kernels + the C ABI (libmctsb200.so; header: include/mctsb200.h)
  _abi.py      ctypes mirror of the header
  _lib.py      loader + Context wrapper (raises if the CUDA library / a B200 is missing: no CPU fallback)
  models.py    SimpleGridWorld / MountainCar parameter structs (device plugins live in csrc/models.cuh), PluginMDP
  plugins/     MDP plugins registered at run time (CUDA text compiled with NVRTC around the tile kernel)
  solvers.py   MCTSSolver / DPWSolver / solve / action / action_info / value / ranked_actions mirror
  distributed.py  root sharding across ranks and the root-parallel allreduce
  workloads.py    the BASELINE.json configurations as synthetic workloads (root formulas, models, solver settings)
  julia/       the Julia glue a maintainer would add to MCTS.jl (unexecuted here: no Julia in the image)
"""
from . import _abi as abi
from ._lib import Context, MCTSB200Error, build_library, load_library, register_mdp
from .models import ContinuousPendulum, GaussianBox, GWPos, InvertedPendulum, MountainCar, MountainCarPlugin, NoisyMountainCar, PluginMDP, SimpleGridWorld
from .solvers import (ConstantPolicy, DPWPlanner, DPWSolver, EnergyPolicy, ExceptionRethrow, FunctionPolicy, MCTSPlanner, MCTSSolver, PluginPolicy,
                      RandomPolicy, RandomSolver, RolloutEstimator, RolloutSimulator, SeekTarget, build_cfg, clear_tree, ranked_actions,
                      simulate, simulate_batch, solve)

__all__ = [
    "abi", "Context", "MCTSB200Error", "build_library", "load_library", "ContinuousPendulum", "GaussianBox", "GWPos", "InvertedPendulum", "MountainCar", "SimpleGridWorld", "register_mdp", "PluginMDP", "MountainCarPlugin", "NoisyMountainCar", "PluginPolicy",
    "ConstantPolicy", "FunctionPolicy", "DPWPlanner", "DPWSolver", "EnergyPolicy", "ExceptionRethrow", "MCTSPlanner", "MCTSSolver",
    "RandomPolicy", "RandomSolver", "RolloutEstimator", "RolloutSimulator", "SeekTarget", "build_cfg", "clear_tree", "ranked_actions",
    "simulate", "simulate_batch", "solve",
]
