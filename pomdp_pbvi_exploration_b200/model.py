"""
POMDP model container -- host-side tensor construction + device-resident tables for the CUDA engine.

Mirrors the constructor contract and attribute names of the reference's `pomdp.Model`
(reference: src/pomdp.py:44-308 on top of src/mdp.py:52-590) so that user code building a model for
the reference builds the same model here.  The tensors the kernels consume are

    reachable_states                          [S,A,R] int64   (src/mdp.py:296-353)
    reachable_probabilities                   [S,A,R] f64     (src/mdp.py:341-353)
    observation_table                         [S,A,O] f64     indexed by the LANDING state
    reachable_transitional_observation_table  [S,A,O,R] f64   (src/pomdp.py:197-205)
    expected_rewards_table                    [S,A] f64       (src/pomdp.py:231-254)

They are built once on the host (one-off, ms..s) and mirrored to the GPU in a kernel-friendly
layout by `pbvi_model_create` (include/pbvi_b200.h); the handle is created lazily on first device
use.  There is no CPU compute path: `use_gpu` style switches of the reference are accepted but the
engine always runs on the device.
"""
from __future__ import annotations

import pickle
import os
import random
from datetime import datetime
from inspect import signature
from typing import Union

import numpy as np


def log(content: str) -> None:
    """Timestamped print, same format as the reference's `log` (src/mdp.py:40-49)."""
    print(f'[{datetime.now().strftime("%m/%d/%Y, %H:%M:%S")}] ' + content)


_VERBOSE = os.environ.get("PBVI_B200_VERBOSE", "0") not in ("0", "", "false")


def _vlog(msg: str) -> None:
    if _VERBOSE:
        log(msg)


class Model:
    """
    POMDP model.  Parameters and attributes follow the reference (src/pomdp.py:147-160 and the attribute
    list in its docstring, src/pomdp.py:82-145).

    Parameters
    ----------
    states : int | list[str] | list[list[str]]
    actions : int | list
    observations : int | list
    transitions : array [S,A,S] | callable(s,a,s_p) | None
    reachable_states : array [S,A,R] | None
    rewards : array [S,A,S,O] | callable(s,a,s_p,o) | None
    observation_table : array [S,A,O] | None
    rewards_are_probabilistic : bool
    state_grid : list[list[int|str]] | None
    start_probabilities : array [S] | None
    end_states, end_actions : list[int]
    """

    def __init__(self,
                 states: Union[int, list],
                 actions: Union[int, list],
                 observations: Union[int, list],
                 transitions=None,
                 reachable_states=None,
                 rewards=None,
                 observation_table=None,
                 rewards_are_probabilistic: bool = False,
                 state_grid=None,
                 start_probabilities=None,
                 end_states: list = [],
                 end_actions: list = []):
        self._device_handle = None
        self.is_on_gpu = True      # the engine is always device-backed; kept for API compatibility

        # ---- label spaces (src/mdp.py:159-192, src/pomdp.py:176-182)
        self.state_grid = None
        if isinstance(states, int):
            self.state_labels = [f's_{i}' for i in range(states)]
        elif isinstance(states, list) and len(states) > 0 and all(isinstance(row, list) for row in states):
            width = len(states[0])
            assert all(len(row) == width for row in states), "All sublists of states must be of equal size"
            self.state_labels = [lab for row in states for lab in row]
            self.state_grid = np.arange(len(states) * width).reshape(len(states), width)
        else:
            self.state_labels = [lab for lab in states if isinstance(lab, str)]
        self.state_count = len(self.state_labels)
        self.states = np.arange(self.state_count)

        self.action_labels = [f'a_{i}' for i in range(actions)] if isinstance(actions, int) else actions
        self.action_count = len(self.action_labels)
        self.actions = np.arange(self.action_count)

        self.observation_labels = [f'o_{i}' for i in range(observations)] if isinstance(observations, int) else observations
        self.observation_count = len(self.observation_labels)
        self.observations = np.arange(self.observation_count)
        S, A, O = self.state_count, self.action_count, self.observation_count
        _vlog(f'POMDP model: {S} states, {A} actions, {O} observations')

        # ---- transitions (src/mdp.py:194-241)
        self.reachable_states = None
        if reachable_states is not None:
            self.reachable_states = np.array(reachable_states)
            assert self.reachable_states.shape[:2] == (S, A), \
                f"Reachable states provided is not of the expected shape (received {self.reachable_states.shape}, expected ({S}, {A}, :))"
            self.reachable_state_count = self.reachable_states.shape[2]

        self.transition_table = None
        self.transition_function = None
        if transitions is None:
            if reachable_states is None:
                rnd = np.random.rand(S, A, S)
                self.transition_table = rnd / np.sum(rnd, axis=2, keepdims=True)
        elif callable(transitions):
            self.transition_function = transitions
            try:
                self.transition_table = np.fromfunction(transitions, (S, A, S))
            except MemoryError:
                self.transition_table = None
        else:
            self.transition_table = np.array(transitions)
            assert self.transition_table.shape == (S, A, S), \
                f"Transitions table provided doesnt have the right shape, it should be SxAxS (expected {(S, A, S)}, received {self.transition_table.shape})"

        self.rewards_are_probabilistic = rewards_are_probabilistic

        # ---- state grid (src/mdp.py:246-282)
        if state_grid is None and self.state_grid is None:
            self.state_grid = np.arange(S).reshape((1, S))
        elif state_grid is not None:
            assert all(isinstance(row, list) for row in state_grid), "The provided states grid must be a list of lists."
            gw = len(state_grid[0])
            assert all(len(row) == gw for row in state_grid), "All rows must have the same length."
            grid = np.zeros((len(state_grid), gw), dtype=int)
            for i, row in enumerate(state_grid):
                for j, el in enumerate(row):
                    if isinstance(el, str):
                        assert el in self.state_labels, f"Countains a state ('{el}') not in the list of states..."
                        grid[i, j] = self.state_labels.index(el)
                    else:
                        assert int(el) < S, f"Countains a state ('{el}') not in the list of states..."
                        grid[i, j] = int(el)
            self.state_grid = grid

        # ---- start distribution, terminal conditions (src/mdp.py:284-294)
        if start_probabilities is not None:
            assert len(start_probabilities) == S
            self.start_probabilities = np.array(start_probabilities, dtype=float)
        else:
            self.start_probabilities = np.full((S,), 1 / S)
        self.end_states = end_states
        self.end_actions = end_actions

        # ---- reachable-state table derived from T when not supplied (src/mdp.py:296-339)
        if self.reachable_states is None:
            self.reachable_states = self._derive_reachable_states()
            self.reachable_state_count = self.reachable_states.shape[2]
        R = self.reachable_state_count

        # ---- reachable probabilities (src/mdp.py:341-353)
        if self.transition_function is None and self.transition_table is None:
            self.reachable_probabilities = np.full(self.reachable_states.shape, 1 / R)
        elif self.transition_table is not None:
            self.reachable_probabilities = self.transition_table[self.states[:, None, None], self.actions[None, :, None], self.reachable_states]
        else:
            rs = self.reachable_states
            self.reachable_probabilities = np.fromfunction(
                lambda s, a, ri: self.transition_function(s.astype(int), a.astype(int), rs[s.astype(int), a.astype(int), ri.astype(int)]),
                rs.shape)

        # ---- observations (src/pomdp.py:184-193)
        if observation_table is None:
            rnd = np.random.rand(S, A, O)
            self.observation_table = rnd / np.sum(rnd, axis=2, keepdims=True)
        else:
            self.observation_table = np.array(observation_table)
            assert self.observation_table.shape == (S, A, O), \
                f"Observations table doesnt have the right shape, it should be SxAxO (expected: {(S, A, O)}, received: {self.observation_table.shape})."

        # ---- RTO[s,a,o,r] = P[s,a,r] * Obs[landing(s,a,r), a, o] (src/pomdp.py:201-202)
        landing_obs = self.observation_table[self.reachable_states[:, :, None, :],
                                             self.actions[None, :, None, None],
                                             self.observations[None, None, :, None]]
        self.reachable_transitional_observation_table = np.einsum('sar,saor->saor', self.reachable_probabilities, landing_obs)

        # ---- rewards (src/pomdp.py:207-254)
        self.immediate_reward_table = None
        self.immediate_reward_function = None
        if rewards is None:
            if len(self.end_states) > 0 or len(self.end_actions) > 0:
                self.immediate_reward_function = self._end_reward_function
            else:
                self.immediate_reward_table = np.random.rand(S, A, S, O)
        elif callable(rewards):
            assert len(signature(rewards).parameters) == 4, "Reward function should accept 4 parameters: s, a, sn, o..."
            self.immediate_reward_function = rewards
        else:
            self.immediate_reward_table = np.array(rewards)
            assert self.immediate_reward_table.shape == (S, A, S, O), \
                f"Rewards table doesnt have the right shape, it should be SxAxSxO (expected: {(S, A, S, O)}, received {self.immediate_reward_table.shape})"

        if self.immediate_reward_table is not None:
            reach_rew = self.immediate_reward_table[self.states[:, None, None, None], self.actions[None, :, None, None],
                                                    self.reachable_states[:, :, :, None], self.observations[None, None, None, :]]
        else:
            rs = self.reachable_states
            fn = self.immediate_reward_function
            reach_rew = np.fromfunction(
                lambda s, a, ri, o: fn(s.astype(int), a.astype(int), rs[s.astype(int), a.astype(int), ri.astype(int)], o.astype(int)),
                (*rs.shape, O))
        self._min_reward = float(np.min(reach_rew))
        self._max_reward = float(np.max(reach_rew))
        self.expected_rewards_table = np.einsum('saor,saro->sa', self.reachable_transitional_observation_table, reach_rew)

    # ------------------------------------------------------------------------------------------
    def _derive_reachable_states(self) -> np.ndarray:
        """argwhere(T>0) rows padded with the smallest unused state ids (src/mdp.py:306-335)."""
        S, A = self.state_count, self.action_count
        rows = []
        for s in range(S):
            per_a = []
            for a in range(A):
                if self.transition_table is not None:
                    per_a.append(np.flatnonzero(self.transition_table[s, a, :] > 0).tolist())
                else:
                    per_a.append([sn for sn in range(S) if self.transition_function(s, a, sn) > 0])
            rows.append(per_a)
        R = max(len(l) for per_a in rows for l in per_a)
        for per_a in rows:
            for l in per_a:
                cand = 0
                while len(l) < R:
                    if cand not in l:
                        l.append(cand)
                    cand += 1
        return np.array(rows, dtype=int)

    def _end_reward_function(self, s, a, sn, o):
        return (np.isin(sn, self.end_states) | np.isin(a, self.end_actions)).astype(int)

    # ---- sampling (src/mdp.py:415-438, src/pomdp.py:261-308); host RNG, same draw order as the reference
    def transition(self, s: int, a: int) -> int:
        if self.reachable_state_count == 1:
            return int(self.reachable_states[s, a, 0])
        return int(np.random.choice(a=self.reachable_states[s, a], size=1, p=self.reachable_probabilities[s, a])[0])

    def observe(self, s_p: int, a: int) -> int:
        """`np.random.choice(observations, size=1, p=obs_table[s_p, a])` without its per-call argument checks (13 us, most of the host
        time of an FSVI trajectory): that call is cdf = p.cumsum(); cdf /= cdf[-1]; searchsorted(cdf, random_sample(), 'right') -- the
        cdf rows are tabulated once (same additions, same division), the one uniform is drawn here, so the legacy RNG stream and the
        result are those of the reference (tests: seeded expansions / solves on tiger equal the reference's)."""
        cdf = self.__dict__.get('_obs_cdf')
        if cdf is None:
            cdf = np.cumsum(self.observation_table, axis=2)
            cdf /= cdf[:, :, -1:]
            self._obs_cdf = cdf
        u = np.random.random_sample()
        row = cdf[s_p, a]
        o = 0
        for c in row.tolist():
            if c <= u:
                o += 1
        return int(self.observations[min(o, len(row) - 1)])

    def reward(self, s: int, a: int, s_p: int, o: int) -> Union[int, float]:
        r = float(self.immediate_reward_table[s, a, s_p, o] if self.immediate_reward_table is not None
                  else self.immediate_reward_function(s, a, s_p, o))
        if self.rewards_are_probabilistic:
            return 1 if random.random() < r else 0
        return r

    def get_coords(self, item):
        items = [item] if isinstance(item, int) else item
        coords = [np.argwhere(self.state_grid == s)[0] for s in items]
        return coords[0] if isinstance(item, int) else coords

    # ---- persistence (src/mdp.py:488-530): pickle of the host tensors (the device handle is rebuilt lazily)
    def __getstate__(self):
        d = dict(self.__dict__)
        d['_device_handle'] = None
        d['_start_device'] = None
        d.pop('_obs_cdf', None)
        return d

    def save(self, file_name: str, path: str = './Models') -> None:
        os.makedirs(path, exist_ok=True)
        if not file_name.endswith('.pck'):
            file_name += '.pck'
        with open(path + '/' + file_name, 'wb') as f:
            pickle.dump(self, f)

    @classmethod
    def load_from_file(cls, file: str) -> 'Model':
        with open(file, 'rb') as f:
            return pickle.load(f)

    # ---- device mirror
    @property
    def gpu_model(self) -> 'Model':
        return self

    @property
    def cpu_model(self) -> 'Model':
        return self

    @property
    def start_belief_device(self):
        """start_probabilities as a cached CUDA tensor (the initial belief b0)."""
        if getattr(self, '_start_device', None) is None:
            import torch
            self._start_device = torch.as_tensor(self.start_probabilities, dtype=torch.float64).to(self.device.device)
        return self._start_device

    @property
    def device(self):
        """The native model handle (`DeviceModel`), created on first use.  Raises if CUDA/the library is missing."""
        if self._device_handle is None:
            from ._native import DeviceModel
            self._device_handle = DeviceModel(self.reachable_states, self.reachable_probabilities,
                                              self.reachable_transitional_observation_table, self.expected_rewards_table)
        return self._device_handle
