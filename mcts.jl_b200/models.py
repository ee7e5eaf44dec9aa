"""Host-side descriptions of the MDPs that have a device plugin (mcts.jl_b200/csrc/models.cuh).

These mirror the POMDPModels.jl constructors the reference's tests and benches use
(`SimpleGridWorld()` everywhere in /root/reference/test, `GWPos`; MountainCar per BASELINE.json) and
only carry parameters: all model arithmetic runs on the device.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field
from typing import Any, Dict, Iterable, Optional, Sequence, Set, Tuple

import numpy as np

from . import _abi as abi


def GWPos(x: int, y: int) -> Tuple[int, int]:
    """SVector{2,Int} stand-in."""
    return (int(x), int(y))


_DEFAULT_REWARDS = {(4, 3): -10.0, (4, 6): -5.0, (9, 3): 10.0, (8, 8): 3.0}


@dataclass
class SimpleGridWorld:
    """POMDPModels.SimpleGridWorld: size, rewards, terminate_from, tprob, discount."""

    size: Tuple[int, int] = (10, 10)
    rewards: Dict[Tuple[int, int], float] = field(default_factory=lambda: dict(_DEFAULT_REWARDS))
    terminate_from: Optional[Set[Tuple[int, int]]] = None  # default: the reward cells
    tprob: float = 0.7
    discount: float = 0.95

    name = "SimpleGridWorld"
    mdp_id = abi.MDP_GRIDWORLD
    actions: Tuple[str, ...] = ("up", "down", "left", "right")
    state_dtype = np.int64
    state_width = 2
    dense = True

    def __post_init__(self):
        if self.terminate_from is None:
            self.terminate_from = set(self.rewards.keys())

    # -- POMDPs.jl-style accessors used by host code --
    def isterminal(self, s) -> bool:
        return s[0] < 0 or s[1] < 0

    def n_states(self) -> int:
        return self.size[0] * self.size[1] + 1

    def state_index(self, s) -> int:
        if self.isterminal(s):
            return self.size[0] * self.size[1]
        return (s[0] - 1) + (s[1] - 1) * self.size[0]

    def state_from_index(self, i: int) -> Tuple[int, int]:
        if i == self.size[0] * self.size[1]:
            return (-1, -1)
        return (i % self.size[0] + 1, i // self.size[0] + 1)

    def states(self) -> Iterable[Tuple[int, int]]:
        return (self.state_from_index(i) for i in range(self.n_states()))

    def action_index(self, a) -> int:
        return self.actions.index(a)

    def encode_states(self, states: Sequence) -> np.ndarray:
        arr = np.asarray(states, dtype=np.int64).reshape(-1, 2)
        return np.ascontiguousarray(arr)

    def decode_state(self, row) -> Tuple[int, int]:
        return (int(row[0]), int(row[1]))

    def params(self) -> abi.GridWorldParams:
        p = abi.GridWorldParams()
        p.nx, p.ny = int(self.size[0]), int(self.size[1])
        p.tprob, p.discount = float(self.tprob), float(self.discount)
        if len(self.rewards) > abi.GW_MAX_SPECIAL or len(self.terminate_from) > abi.GW_MAX_SPECIAL:
            raise ValueError(f"at most {abi.GW_MAX_SPECIAL} reward / terminate_from cells are supported")
        p.n_rewards = len(self.rewards)
        for i, ((x, y), v) in enumerate(self.rewards.items()):
            p.reward_x[i], p.reward_y[i], p.reward_v[i] = int(x), int(y), float(v)
        p.n_terminate = len(self.terminate_from)
        for i, (x, y) in enumerate(sorted(self.terminate_from)):
            p.term_x[i], p.term_y[i] = int(x), int(y)
        return p


@dataclass
class MountainCar:
    """POMDPModels.MountainCar: state (x, v), actions [-1., 0., 1.]."""

    discount: float = 0.99
    cost: float = -1.0
    jackpot: float = 100.0

    name = "MountainCar"
    mdp_id = abi.MDP_MOUNTAINCAR
    actions: Tuple[float, ...] = (-1.0, 0.0, 1.0)
    state_dtype = np.float64
    state_width = 2
    dense = False

    def isterminal(self, s) -> bool:
        return s[0] >= 0.5

    def action_index(self, a) -> int:
        return self.actions.index(float(a))

    def encode_states(self, states: Sequence) -> np.ndarray:
        return np.ascontiguousarray(np.asarray(states, dtype=np.float64).reshape(-1, 2))

    def decode_state(self, row) -> Tuple[float, float]:
        return (float(row[0]), float(row[1]))

    def params(self) -> abi.MountainCarParams:
        return abi.MountainCarParams(float(self.discount), float(self.cost), float(self.jackpot))


@dataclass
class GaussianBox:
    """The continuous-action MDP of the reference's test/discussion_514.jl -- states(m) = Box, actions(m) = Box(a_lo, a_hi),
    transition(m, s, a) = MvNormal(s, Diagonal(var)), reward 0, discount 0.9 -- with an optional drift (sp = s + drift * a + noise)
    and quadratic cost (r = -cost * |sp|^2) that give a search something to find (include/mctsb200.h:
    mctsb200_gaussian_box_params). States are 3 coordinates (a fourth, carried along unchanged, pads the ABI state); actions are
    2-vectors, so `action` returns a tuple of floats and only DPWSolver can plan."""

    a_lo: Tuple[float, float] = (-5.0, -5.0)
    a_hi: Tuple[float, float] = (5.0, 5.0)
    var: Tuple[float, float, float] = (0.1, 0.1, 0.1)
    drift: float = 0.0
    cost: float = 0.0
    discount: float = 0.9

    name = "GaussianBox"
    mdp_id = abi.MDP_GAUSSIAN_BOX
    actions = None          # a continuum: Box(a_lo, a_hi)
    action_width = abi.BOX_ACTION_DOUBLES
    state_dtype = np.float64
    state_width = abi.BOX_STATE_DOUBLES
    dense = False

    def isterminal(self, s) -> bool:
        return False

    def action_in_space(self, a) -> bool:
        """`a in actions(m)` (test/discussion_514.jl:18)."""
        return len(a) == self.action_width and all(lo <= x <= hi for x, lo, hi in zip(a, self.a_lo, self.a_hi))

    def encode_states(self, states: Sequence) -> np.ndarray:
        arr = np.zeros((len(states), self.state_width), dtype=np.float64)
        for i, s in enumerate(states):
            arr[i, : len(s)] = s
        return arr

    def decode_state(self, row) -> Tuple[float, ...]:
        return tuple(float(x) for x in row[:3])

    def params(self) -> abi.GaussianBoxParams:
        p = abi.GaussianBoxParams()
        p.discount = float(self.discount)
        for i in range(abi.BOX_ACTION_DOUBLES):
            p.a_lo[i], p.a_hi[i] = float(self.a_lo[i]), float(self.a_hi[i])
        for i in range(3):
            p.var[i] = float(self.var[i])
        p.drift, p.cost = float(self.drift), float(self.cost)
        return p


PLUGIN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "plugins")


@dataclass
class PluginMDP:
    """An MDP whose device plugin is registered at run time (include/mctsb200.h: mctsb200_mdp_register) -- the stand-in for
    defining POMDPs.gen / isterminal / actions / discount for a new model in Julia. `source` is CUDA text defining
    `struct <struct_name>` (interface: csrc/custom.cuh); `params` is a ctypes.Structure mirroring its `Params` POD
    (`double discount` first). States are tuples of `state_width` float64 values (2 or 4)."""

    name: str
    source: str
    struct_name: str
    actions: Tuple[Any, ...]
    params_struct: C.Structure
    state_width: int = 2
    max_branch: int = 0        # distinct successors of one (s, a): 0 = unbounded (always correct, the default); 1 = deterministic
                               # transitions, an explicit opt-in: the kernels then reuse the first sampled successor
    terminal: Any = None       # optional host-side isterminal(s) (DPW's terminal-root error path in the Python mirror)
    lib_path: Optional[str] = None
    oracle_mdp_id: Optional[int] = None  # tests only: the id under which oracle/mcts_oracle.cpp restates this model
    # a continuous action space (mctsb200_mdp_register_continuous): actions(mdp) = Box(a_lo, a_hi) of action_width coordinates, `actions`
    # is None, `action` returns a tuple of floats and only DPWSolver can plan
    action_width: int = 0
    a_lo: Tuple[float, ...] = ()
    a_hi: Tuple[float, ...] = ()

    state_dtype = np.float64
    dense = False
    _mdp_id: Optional[int] = field(default=None, init=False, repr=False)

    @classmethod
    def from_file(cls, path: str, **kw) -> "PluginMDP":
        with open(path if os.path.isabs(path) else os.path.join(PLUGIN_DIR, path), encoding="utf-8") as f:
            return cls(source=f.read(), **kw)

    @property
    def mdp_id(self) -> int:
        if self._mdp_id is None:
            from ._lib import register_mdp, register_mdp_continuous

            if self.action_width:
                self._mdp_id = register_mdp_continuous(self.name, self.source, self.struct_name, self.state_width, self.action_width, self.lib_path)
            else:
                self._mdp_id = register_mdp(self.name, self.source, self.struct_name, self.state_width, len(self.actions), self.max_branch,
                                            self.lib_path)
        return self._mdp_id

    def action_in_space(self, a) -> bool:
        """`a in actions(m)` for a continuous action space (test/discussion_514.jl:18)."""
        return len(a) == self.action_width and all(lo <= x <= hi for x, lo, hi in zip(a, self.a_lo, self.a_hi))

    @property
    def discount(self) -> float:
        return float(self.params_struct.discount)

    def isterminal(self, s) -> bool:
        return bool(self.terminal(s)) if self.terminal is not None else False

    def action_index(self, a) -> int:
        return self.actions.index(a)

    def encode_states(self, states: Sequence) -> np.ndarray:
        return np.ascontiguousarray(np.asarray(states, dtype=np.float64).reshape(-1, self.state_width))

    def decode_state(self, row) -> Tuple[float, ...]:
        return tuple(float(x) for x in row)

    def params(self) -> C.Structure:
        return self.params_struct


class ContinuousPendulumParams(C.Structure):
    """Params of plugins/continuous_pendulum.cuh."""

    _fields_ = [("discount", C.c_double), ("g", C.c_double), ("m", C.c_double), ("l", C.c_double), ("M", C.c_double), ("alpha", C.c_double),
                ("dt", C.c_double), ("cost", C.c_double), ("u_max", C.c_double), ("noise", C.c_double), ("effort", C.c_double), ("gain", C.c_double)]


def ContinuousPendulum(u_max: float = 50.0, noise: float = 10.0, effort: float = 0.0, gain: float = 100.0, g: float = 9.81, m: float = 2.0,
                       l: float = 8.0, M: float = 8.0, dt: float = 0.1, discount: float = 0.99, cost: float = -1.0,
                       name: str = "ContinuousPendulum") -> "PluginMDP":
    """An inverted pendulum with a continuous torque (plugins/continuous_pendulum.cuh): a run-time plugin whose action space is
    Box([-u_max], [u_max]) -- the reference's test/discussion_514.jl case (actions(m) = Box) for a model the caller brings. State
    (theta, omega); `action` returns a 1-tuple; rollout policy id 100 is a proportional controller."""
    p = ContinuousPendulumParams(float(discount), float(g), float(m), float(l), float(M), 1.0 / (float(m) + float(M)), float(dt), float(cost),
                                 float(u_max), float(noise), float(effort), float(gain))
    return PluginMDP.from_file("continuous_pendulum.cuh", name=name, struct_name="ContinuousPendulum", actions=None, params_struct=p,
                               state_width=2, max_branch=0, terminal=lambda s: abs(s[0]) > 1.5707963267948966, oracle_mdp_id=1002,
                               action_width=1, a_lo=(-float(u_max),), a_hi=(float(u_max),))


class NoisyMountainCarParams(C.Structure):
    """Params of plugins/noisy_mountaincar.cuh."""

    _fields_ = [("discount", C.c_double), ("cost", C.c_double), ("jackpot", C.c_double), ("noise", C.c_double),
                ("push", C.c_double * 4), ("n_actions", C.c_int32), ("pad", C.c_int32)]


def MountainCarPlugin(discount: float = 0.99, cost: float = -1.0, jackpot: float = 100.0, name: str = "MountainCarPlugin") -> PluginMDP:
    """The built-in MountainCar written as a run-time plugin (plugins/mountaincar.cuh): same arithmetic, so searches are
    bit-identical to `MountainCar()` -- the parity anchor of the registration path."""
    return PluginMDP.from_file("mountaincar.cuh", name=name, struct_name="MountainCarPlugin", actions=(-1.0, 0.0, 1.0),
                               params_struct=abi.MountainCarParams(float(discount), float(cost), float(jackpot)), state_width=2, max_branch=1,
                               terminal=lambda s: s[0] >= 0.5, oracle_mdp_id=abi.MDP_MOUNTAINCAR)


def NoisyMountainCar(discount: float = 0.99, cost: float = -1.0, jackpot: float = 100.0, noise: float = 0.002,
                     actions: Tuple[float, ...] = (-1.0, 0.0, 1.0), name: Optional[str] = None, state_width: int = 2) -> PluginMDP:
    """MountainCar with uniform velocity noise (plugins/noisy_mountaincar.cuh): a continuous-state MDP with stochastic
    transitions -- every gen call can produce a new state, the case double progressive widening exists for. `actions` are
    2 to 4 accelerations (velocity increment a * 0.001); `state_width=4` registers the same model with two padding
    coordinates (carried along unchanged), which exercises the 4-double host shapes."""
    if not 2 <= len(actions) <= 4:
        raise ValueError("2 to 4 actions")
    p = NoisyMountainCarParams(float(discount), float(cost), float(jackpot), float(noise))
    for i, a in enumerate(actions):
        p.push[i] = float(a) * 0.001
    p.n_actions = len(actions)
    with open(os.path.join(PLUGIN_DIR, "noisy_mountaincar.cuh"), encoding="utf-8") as f:
        src = f.read()
    if state_width == 4:  # the padding coordinates are part of the state (and of its hash key); the dynamics ignore them
        src = src.replace("        sp[0] = x_;\n", "        sp[0] = x_;\n        sp[2] = s[2];\n        sp[3] = s[3];\n")
        assert "sp[3] = s[3]" in src
    elif state_width != 2:
        raise ValueError("state_width 2 or 4")
    return PluginMDP(name=name or f"NoisyMountainCar{state_width}x{len(actions)}", source=src, struct_name="NoisyMountainCar",
                     actions=tuple(float(a) for a in actions), params_struct=p, state_width=state_width, max_branch=0,
                     terminal=lambda s: s[0] >= 0.5, oracle_mdp_id=1000 if state_width == 2 else None)


class InvertedPendulumParams(C.Structure):
    """Params of plugins/inverted_pendulum.cuh."""

    _fields_ = [("discount", C.c_double), ("g", C.c_double), ("m", C.c_double), ("l", C.c_double), ("M", C.c_double), ("alpha", C.c_double),
                ("dt", C.c_double), ("cost", C.c_double), ("force", C.c_double * 3), ("noise", C.c_double)]


def InvertedPendulum(g: float = 9.81, m: float = 2.0, l: float = 8.0, M: float = 8.0, dt: float = 0.1, discount: float = 0.99,
                     cost: float = -1.0, noise: float = 10.0, name: str = "InvertedPendulum") -> PluginMDP:
    """POMDPModels.InvertedPendulum (notebooks/Test_Visualization.ipynb cell 4) as a run-time plugin (plugins/inverted_pendulum.cuh):
    state (theta, omega), actions (-50., 0., 50.), uniform force noise -- continuous stochastic transitions, max_branch = 0."""
    p = InvertedPendulumParams(float(discount), float(g), float(m), float(l), float(M), 1.0 / (float(m) + float(M)), float(dt), float(cost))
    p.force[0], p.force[1], p.force[2] = -50.0, 0.0, 50.0
    p.noise = float(noise)
    return PluginMDP.from_file("inverted_pendulum.cuh", name=name, struct_name="InvertedPendulum", actions=(-50.0, 0.0, 50.0), params_struct=p,
                               state_width=2, max_branch=0, terminal=lambda s: abs(s[0]) > 1.5707963267948966, oracle_mdp_id=1001)
