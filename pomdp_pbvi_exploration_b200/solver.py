"""
Point-Based Value Iteration solver with the reference's interface (src/pomdp.py:1299-2578, src/mdp.py:1414-1525),
running on the B200 engine (libpbvi_b200.so through `_native.DeviceModel`).

Host Python keeps what the reference keeps in Python -- the expand/backup loop, set bookkeeping on 16-byte row keys,
host RNG draws (same generators, same draw order as the reference's CPU path, so seeded runs pick the same states /
actions / observations) -- and every array operation of the hot path is a kernel of the library:

    backup            pbvi_backup_select + pbvi_group_keys (distinct generating tuples) + pbvi_backup_assemble + pbvi_group_keys /
                      pbvi_confirm_groups (the byte-dedup, on the device); host-resident belief sets travel packed
                      (pbvi_pack_slabs_host + pbvi_unpack_rows)
    compute_change    pbvi_max_values
    Belief.update     pbvi_belief_update (bit-identical to the reference, NaN rows included)
    SSEA / GER / HSVI pbvi_belief_successors, pbvi_min_l2_distance, pbvi_ger_scores, pbvi_sawtooth,
                      pbvi_observation_probabilities
    VI_Solver         pbvi_vi_sweep

There is no CPU path: `use_gpu` is accepted for signature compatibility and ignored.
"""
from __future__ import annotations

import random
from datetime import datetime
from typing import Union

import numpy as np
import torch

from .belief import Belief, BeliefSet
from .model import Model, log
from .sets import _dedup_exact_host, dedup_rows, group_by_key, unique_rows_first  # noqa: F401
from .value_function import AlphaVector, ValueFunction  # noqa: F401


def _now_synced(model: Model) -> datetime:
    """Wall clock after the device drained, so the recorded stage times are device-complete (the reference's CuPy path
    reads the clock without synchronising)."""
    torch.cuda.synchronize(model.device.device)
    return datetime.now()


# =====================================================================================================================
class MDPSolverHistory:
    """History of a value-iteration run (reference src/mdp.py:1281-1400, data fields and summary; plots out of scope)."""

    def __init__(self, tracking_level: int, model: Model, gamma: float, eps: float, initial_value_function=None):
        self.tracking_level = tracking_level
        self.model = model
        self.gamma = gamma
        self.eps = eps
        self.run_ts = datetime.now()
        self.iteration_times = []
        self.value_function_changes = []
        self.value_functions = []
        if self.tracking_level >= 2:
            self.value_functions.append(initial_value_function)

    @property
    def solution(self) -> ValueFunction:
        assert self.tracking_level >= 2, "Tracking level is set too low, increase it to 2 if you want to have value function tracking as well."
        return self.value_functions[-1]

    def add(self, iteration_time: float, value_function_change: float, value_function) -> None:
        if self.tracking_level >= 1:
            self.iteration_times.append(float(iteration_time))
            self.value_function_changes.append(float(value_function_change))
        if self.tracking_level >= 2:
            self.value_functions.append(value_function)

    @property
    def summary(self) -> str:
        s = 'Summary of Value Iteration run'
        s += f'\n  - Model: {self.model.state_count}-state, {self.model.action_count}-action'
        s += f'\n  - Converged in {len(self.iteration_times)} iterations and {sum(self.iteration_times):.4f} seconds'
        if self.tracking_level >= 1:
            s += f'\n  - Took on average {sum(self.iteration_times) / len(self.iteration_times):.4f}s per iteration'
        return s


class VI_Solver:
    """
    MDP value iteration (reference src/mdp.py:1414-1525): alpha[a,s] = Rbar[s,a] + gamma * sum_r P[s,a,r] V*[reach[s,a,r]],
    until max |V* - V*_old| < eps * gamma / (1 - gamma).  One `pbvi_vi_sweep` launch per iteration.
    """

    def __init__(self, horizon: int = 10000, gamma: float = 0.99, eps: float = 0.001):
        self.horizon = horizon
        self.gamma = gamma
        self.eps = eps

    def solve(self, model: Model, initial_value_function: Union[ValueFunction, None] = None, use_gpu: bool = False,
              history_tracking_level: int = 1, print_progress: bool = True):
        dev = model.device
        if initial_value_function is None:
            V = ValueFunction(model, model.expected_rewards_table.T, model.actions)
        else:
            V = initial_value_function
        v_opt = torch.max(V.alpha_vector_array, dim=0).values
        history = MDPSolverHistory(history_tracking_level, model, self.gamma, self.eps, V)
        max_allowed_change = self.eps * (self.gamma / (1 - self.gamma))
        alpha = V.alpha_vector_array
        swept = False
        for _ in range(self.horizon):
            start = datetime.now()
            alpha, v_new = dev.vi_sweep(v_opt, self.gamma)
            swept = True
            max_change = float(torch.max(torch.abs(v_new - v_opt)))          # one scalar D2H per sweep = the convergence test
            v_opt = v_new
            if history_tracking_level >= 2:
                history.add((datetime.now() - start).total_seconds(), max_change, ValueFunction(model, alpha.clone(), model.actions))
            else:
                history.add((datetime.now() - start).total_seconds(), max_change, None)
            if max_change < max_allowed_change:
                break
        if swept:
            V = ValueFunction(model, alpha, model.actions)       # byte-dedup of identical action rows, last action wins
        return V, history


# =====================================================================================================================
class BeliefValueMapping:
    """
    HSVI's upper bound: (belief, value) pairs interpolated with the sawtooth rule over the corner values of an MDP
    solution (reference src/pomdp.py:786-895).  The arrays the interpolation runs over are refreshed by `update()` only (and built
    lazily on first use), as in the reference; a belief that is already stored returns its stored value.  Divergence (documented,
    SURVEY.md section 4): the ratio min_s b(s)/b_i(s) runs over the support of b_i -- the reference's formula takes 0/0 = NaN on
    sparse beliefs, which disables its own HSVI outside tiger-class models; on strictly positive beliefs the two agree.

    Device state: the stored beliefs live in one growing row store with their 128-bit keys and values next to them, and -- for
    the rows the interpolation covers -- as SUPPORT LISTS (`pbvi_support_lists`: states / values of the positive entries and
    b_i . corner), built once per row, so that a level of the exploration (`pbvi_hsvi_level`) touches nnz instead of S entries per
    stored belief and needs no host round trip for the lookups.
    """

    def __init__(self, model: Model, corner_belief_values: ValueFunction) -> None:
        self.model = model
        self.corner_belief_values = corner_belief_values
        self.corner_values = torch.max(corner_belief_values.alpha_vector_array, dim=0).values.contiguous()
        self.beliefs = []
        self.belief_value_mapping = {}
        self._cap = 0
        self._rows = self._keys_dev = self._vals_dev = None
        self._idx = self._val = self._count = self._dot = None
        self._n_listed = 0          # rows that have support lists
        self._n_ub = None           # rows the interpolation covers (None: arrays not built yet -- built lazily on first use)

    # ---- storage ------------------------------------------------------------------------------------------------------
    def _reserve(self, n: int) -> None:
        if n <= self._cap:
            return
        dev, S = self.model.device.device, self.model.state_count
        cap = max(64, 2 * n)
        def grown(old, shape, dtype):
            t = torch.empty(shape, dtype=dtype, device=dev)
            if old is not None and len(self.beliefs):
                t[:old.shape[0]] = old
            return t
        self._rows = grown(self._rows, (cap, S), torch.float64)
        self._keys_dev = grown(self._keys_dev, (cap, 2), torch.int64)
        self._vals_dev = grown(self._vals_dev, (cap,), torch.float64)
        self._idx = grown(self._idx, (cap, S), torch.int32)
        self._val = grown(self._val, (cap, S), torch.float64)
        self._count = grown(self._count, (cap,), torch.int32)
        self._dot = grown(self._dot, (cap,), torch.float64)
        self._cap = cap

    def _key(self, b: Belief):
        k = b.__dict__.get('_row_key')
        if k is None:
            k = b._row_key = tuple(self.model.device.row_hash(b.values[None, :]).cpu().numpy()[0].tolist())
        return k

    def _append(self, b: Belief, key: tuple, v: float, on_device: bool) -> None:
        """Bookkeeping of one new pair; `on_device`: the level kernel already wrote the key and the value."""
        n = len(self.beliefs)
        self._reserve(n + 1)
        self._rows[n] = b.values
        if not on_device:
            self._keys_dev[n] = torch.tensor(key, dtype=torch.int64)
            self._vals_dev[n] = float(v)
        self.beliefs.append(b)
        self.belief_value_mapping[key] = v

    def add(self, b: Belief, v: float) -> None:
        """Adds (belief, value) unless the belief is already stored (reference :838-850).  Identity is the 128-bit row key."""
        k = self._key(b)
        if k not in self.belief_value_mapping:
            self._append(b, k, v, on_device=False)

    def _cover(self, n: int) -> None:
        """Makes the interpolation arrays cover the first n stored beliefs (support lists for the rows that lack them)."""
        if n > self._n_listed:
            lo = self._n_listed
            self.model.device.support_lists(self._rows[lo:n], self.corner_values, self._idx[lo:n], self._val[lo:n], self._count[lo:n],
                                            self._dot[lo:n])
            self._n_listed = n
        self._n_ub = n

    @property
    def belief_array(self) -> torch.Tensor:
        if self._n_ub is None:
            self._cover(len(self.beliefs))
        return self._rows[:self._n_ub]

    @property
    def value_array(self) -> torch.Tensor:
        if self._n_ub is None:
            self._cover(len(self.beliefs))
        return self._vals_dev[:self._n_ub]

    def update(self) -> None:
        if len(self.beliefs) == 0:
            return
        self._cover(len(self.beliefs))

    def _arrays(self):
        """(idx, val, count, dot, values, n_ub) of the interpolation, built lazily like the reference's array properties."""
        if len(self.beliefs) == 0:
            return None, None, None, None, None, 0
        if self._n_ub is None:
            self._cover(len(self.beliefs))
        return self._idx, self._val, self._count, self._dot, self._vals_dev, self._n_ub

    def evaluate_rows(self, rows: torch.Tensor) -> np.ndarray:
        """
        Upper-bound value of every row [n,S] with two launches and one read-back: stored beliefs return their stored value
        (the reference's shortcut, :884-885), the rest go through the sawtooth kernel against the arrays as of the last
        `update()` (the reference refreshes them only there, :866-871).
        """
        dev = self.model.device
        n = rows.shape[0]
        keys = dev.row_hash(rows).cpu().numpy().tolist()
        idx, val, count, dot, vals, n_ub = self._arrays()
        out = dev.sawtooth_lists(self.corner_values, idx, val, count, dot, vals, n_ub, rows).cpu().numpy()
        for i in range(n):
            hit = self.belief_value_mapping.get(tuple(keys[i]))
            if hit is not None:
                out[i] = hit
        return out

    def evaluate(self, belief: Belief) -> float:
        return float(self.evaluate_rows(belief.values[None, :])[0])


# =====================================================================================================================
class SolverHistory:
    """History of a PBVI run (reference src/pomdp.py:898-1117: data fields and `summary`; plots / videos out of scope)."""

    def __init__(self, tracking_level: int, model: Model, gamma: float, eps: float, expand_function: str, expand_append: bool,
                 initial_value_function: ValueFunction, initial_belief_set: BeliefSet):
        self.tracking_level = tracking_level
        self.model = model
        self.gamma = gamma
        self.eps = eps
        self.run_ts = datetime.now()
        self.expand_function = expand_function
        self.expand_append = expand_append
        self.expansion_times = []
        self.backup_times = []
        self.pruning_times = []
        self.alpha_vector_counts = []
        self.beliefs_counts = []
        self.prune_counts = []
        if self.tracking_level >= 1:
            self.alpha_vector_counts.append(len(initial_value_function))
            self.beliefs_counts.append(len(initial_belief_set))
        self.belief_sets = []
        self.value_functions = []
        self.value_function_changes = []
        if self.tracking_level >= 2:
            self.belief_sets.append(initial_belief_set)
            self.value_functions.append(initial_value_function)

    @property
    def solution(self) -> ValueFunction:
        assert self.tracking_level >= 2, "Tracking level is set too low, increase it to 2 if you want to have value function tracking as well."
        return self.value_functions[-1]

    @property
    def explored_beliefs(self) -> BeliefSet:
        assert self.tracking_level >= 2, "Tracking level is set too low, increase it to 2 if you want to have belief sets tracking as well."
        return self.belief_sets[-1]

    def add_expand_step(self, expansion_time: float, belief_set: BeliefSet) -> None:
        if self.tracking_level >= 1:
            self.expansion_times.append(float(expansion_time))
            self.beliefs_counts.append(len(belief_set))
        if self.tracking_level >= 2:
            self.belief_sets.append(belief_set)

    def add_backup_step(self, backup_time: float, value_function_change: float, value_function: ValueFunction) -> None:
        if self.tracking_level >= 1:
            self.backup_times.append(float(backup_time))
            self.alpha_vector_counts.append(len(value_function))
            self.value_function_changes.append(float(value_function_change))
        if self.tracking_level >= 2:
            self.value_functions.append(value_function)

    def add_prune_step(self, prune_time: float, alpha_vectors_pruned: int) -> None:
        if self.tracking_level >= 1:
            self.pruning_times.append(prune_time)
            self.prune_counts.append(alpha_vectors_pruned)

    @property
    def summary(self) -> str:
        s = 'Summary of Value Iteration run'
        s += f'\n  - Model: {self.model.state_count} state, {self.model.action_count} action, {self.model.observation_count} observations'
        s += f'\n  - Converged or stopped after {len(self.expansion_times)} expansion steps and {len(self.backup_times)} backup steps.'
        if self.tracking_level >= 1:
            s += f'\n  - Resulting value function has {self.alpha_vector_counts[-1]} alpha vectors.'
            s += f'\n  - Converged in {(sum(self.expansion_times) + sum(self.backup_times)):.4f}s'
            s += '\n'
            s += f'\n  - Expand function took on average {sum(self.expansion_times) / len(self.expansion_times):.4f}s '
            if self.expand_append:
                s += f'and yielded on average {sum(np.diff(self.beliefs_counts)) / len(self.beliefs_counts[1:]):.2f} beliefs per iteration.'
            else:
                s += f'and yielded on average {sum(self.beliefs_counts[1:]) / len(self.beliefs_counts[1:]):.2f} beliefs per iteration.'
            s += f' ({np.sum(np.divide(self.expansion_times, self.beliefs_counts[1:])) / len(self.expansion_times):.4f}s/it/belief)'
            s += f'\n  - Backup function took on average {sum(self.backup_times) / len(self.backup_times):.4f}s '
            s += f'and yielded on average {np.average(np.diff(self.alpha_vector_counts)):.2f} alpha vectors per iteration.'
            s += f' ({np.sum(np.divide(self.backup_times, self.alpha_vector_counts[1:])) / len(self.backup_times):.4f}s/it/alpha)'
            s += f'\n  - Pruning function took on average {sum(self.pruning_times) / len(self.pruning_times):.4f}s '
            s += f'and yielded on average prunings of {sum(self.prune_counts) / len(self.prune_counts):.2f} alpha vectors per iteration.'
        return s


# =====================================================================================================================
class PBVI_Solver:
    """
    Point-Based Value Iteration (reference src/pomdp.py:1299-2413).

    Parameters
    ----------
    gamma : float, default=0.99
    eps : float, default=0.001
    expand_function : str, default='ssea'
        One of ra, ssra, ssga, ssea, ger, hsvi, fsvi, fsvi_eg, perseus (matched by substring like the reference).
    expand_function_params
        Extra parameters of the expand function (epsilon, mdp_policy, eps_greedy).
    """

    def __init__(self, gamma: float = 0.99, eps: float = 0.001, expand_function: str = 'ssea', **expand_function_params):
        self.gamma = gamma
        self.eps = eps
        self.expand_function = expand_function
        self.expand_function_params = expand_function_params

    # ------------------------------------------------------------------------------------------------------------
    def test_n_simulations(self, model: Model, value_function: ValueFunction, n: int = 1000, horizon: int = 300, print_progress: bool = False):
        """
        Evaluates a value function with n simultaneous simulations (reference src/pomdp.py:1338-1444): returns the start states,
        the step at which each simulation reached an end state (-1: never), and the per-step rewards / discounted rewards.
        Same host RNG draws as the reference's CPU path; argmax(B.V^T) and the belief update run on the device.
        """
        from .simulation import SimulationSet
        dev = model.device
        beliefs = Belief(model).values[None, :].repeat(n, 1)
        sims = SimulationSet(model)
        start_states = sims.initialize_simulations(n, None)
        sim_is_done = np.zeros(n, dtype=bool)
        done_at_step = np.full(n, -1)
        discount = self.gamma
        rewards, discounted_rewards = [], []
        for i in range(horizon):
            _, best = dev.max_values(beliefs, value_function.alpha_vector_array)
            best_actions = value_function.actions[best.cpu().numpy()]
            sims.is_done = np.zeros(n, dtype=bool)          # the reference keeps stepping finished simulations and masks their rewards
            step_rewards, observations = sims.run_actions(best_actions)
            beliefs, _ = dev.belief_update(beliefs, best_actions.astype(np.int32), observations.astype(np.int32))
            rewards.append(np.where(~sim_is_done, step_rewards, 0))
            discounted_rewards.append(np.where(~sim_is_done, step_rewards * discount, 0))
            are_done = np.isin(sims.agent_states, np.array(model.end_states))
            done_at_step[sim_is_done ^ are_done] = i + 1
            sim_is_done |= are_done
            discount *= self.gamma
            if np.all(sim_is_done):
                break
        return start_states, done_at_step, rewards, discounted_rewards

    # ------------------------------------------------------------------------------------------------------------
    def backup(self, model: Model, belief_set: BeliefSet, value_function: ValueFunction, append: bool = False,
               belief_dominance_prune: bool = True) -> ValueFunction:
        """
        Point-based backup (reference src/pomdp.py:1447-1524).  For every belief: v*[a,o] = argmax_v b.Gamma[a,o,v],
        a* = argmax_a b.(Rbar[:,a] + sum_o Gamma[a,o,v*]), alpha_b = that vector; then the byte-dedup of the
        ValueFunction constructor and, with `append`, the union with the old value function.

        Two beliefs that select the same (a*, v*[a*,:]) tuple produce the same bytes, so only the distinct tuples are
        assembled; the byte-dedup then runs over those rows (different tuples can still give identical rows).
        """
        dev = model.device
        if (not belief_dominance_prune and self.SMALL_PATH and belief_set._device is not None and len(belief_set) > 0 and
                dev.backup_small_eligible(len(belief_set), len(value_function))):
            # tiger / 4x4-class problems: one kernel, one synchronisation (`pbvi_backup_small`) instead of the ~20-launch pipeline
            rows, actions, keys = dev.backup_small(belief_set.belief_array, value_function.alpha_vector_array, self.gamma)
            new_vf = ValueFunction(model, rows, actions, _trusted=True, _hashes=keys)
        else:
            tuples, _, last = self.select_tuples_device(model, belief_set, value_function, belief_dominance_prune, early_rows=True)
            new_vf = self.rows_from_tuples(model, value_function, tuples, last)
        if append:
            n_new = len(new_vf)
            new_vf.extend(value_function)
            # the union keeps the new rows first and every old row (bytewise): max over it = max(new rows, old value function)
            new_vf.parent_uid, new_vf.n_new = value_function.uid, n_new
        return new_vf

    def select_tuples_device(self, model: Model, belief_set: BeliefSet, value_function: ValueFunction, belief_dominance_prune: bool = False,
                             early_rows: bool = False):
        """
        First half of the backup: for every belief the tuple (a*, v*[a*, 0..O-1]) that generates its alpha row, reduced to the
        DISTINCT tuples in order of first occurrence -- on the device (`pbvi_group_keys`), nothing but the count comes back.
        Returns CUDA int32 tensors (tuples [u, 1+O], first [u], last [u]) where first / last are the positions (in this belief
        set, after the optional dominance filter) of the first / last belief that chose the tuple.  A tuple is 4*(1+O) bytes,
        the row it generates 8*S: the sharded backup exchanges these.  `early_rows`: the caller will hand the returned tuples to
        `rows_from_tuples` as they are (the single-process backup), so a streamed select may assemble and read back the rows of the
        tuples it knows before its last chunk (`_select_streamed`).
        """
        dev = model.device
        V = value_function.alpha_vector_array
        nB = len(belief_set)
        if nB == 0:
            z = torch.zeros((0,), dtype=torch.int32, device=dev.device)
            return torch.zeros((0, 1 + dev.O), dtype=torch.int32, device=dev.device), z, z
        # value[b][a] itself is only read by the dominance filter; without it a* alone is asked for and beliefs with a single
        # possible winner skip the exact reference-order sum
        self._early = early = None
        if belief_set._device is None and nB >= 2 * self.STREAM_FIRST_CHUNK:
            if early_rows and self.EARLY_ROWS and not belief_dominance_prune:
                early = {}
            vstar, value, astar = self._select_streamed(model, belief_set, V, want_value=belief_dominance_prune, early=early)
        else:
            vstar, value, astar = dev.backup_select(belief_set.belief_array, V, self.gamma, want_value=belief_dominance_prune)
        B = belief_set.belief_array
        ar = torch.arange(nB, device=dev.device)
        sel = vstar[ar, astar.long()]                                                    # [nB, O]
        keys = torch.cat([astar[:, None], sel], dim=1)
        if belief_dominance_prune:
            # keep b iff b.alpha_b > max_v b.alpha_v, strict (reference :1509-1515)
            best_old, _ = dev.max_values(B, V)
            keep = value[ar, astar.long()] > best_old
            keys = keys[torch.nonzero(keep)[:, 0]]
        first, last, _ = dev.group_keys(keys)
        tuples = keys[first.long()]
        if early:
            # first-occurrence order: the tuples of the rows selected before the last chunk are a prefix of all tuples
            u1 = early['tuples'].shape[0]
            if tuples.shape[0] >= u1 and torch.equal(tuples[:u1], early['tuples']):
                early['token'] = tuples
                self._early = early
        return tuples, first, last

    def select_tuples(self, model: Model, belief_set: BeliefSet, value_function: ValueFunction, belief_dominance_prune: bool = False):
        """`select_tuples_device` as host int64 arrays (tuples [u, 1+O], first [u], last [u])."""
        t, f, l = self.select_tuples_device(model, belief_set, value_function, belief_dominance_prune)
        return tuple(x.cpu().numpy().astype(np.int64) for x in (t, f, l))

    def rows_from_tuples(self, model: Model, value_function: ValueFunction, tuples, last) -> ValueFunction:
        """
        Second half of the backup: assembles the alpha row of every distinct tuple (reference operation order, bit-identical for
        R = 1) and applies the ValueFunction constructor's byte-dedup on the device: rows are grouped by their 128-bit keys
        (accumulated by the assemble kernel), every key match is confirmed bytewise.  Different tuples can generate identical
        bytes; a byte group keeps the position of its first tuple and the action of the tuple whose LAST belief comes latest
        (`last`).  `tuples` [u, 1+O] / `last` [u]: CUDA tensors (from `select_tuples_device`) or host arrays.
        """
        dev = model.device
        t = torch.as_tensor(tuples).to(device=dev.device, dtype=torch.int32)
        n = t.shape[0]
        if n == 0:
            return ValueFunction(model, torch.empty((0, dev.S), dtype=torch.float64, device=dev.device), np.zeros(0, dtype=np.int64))
        rank = torch.as_tensor(last).to(device=dev.device, dtype=torch.int32)
        early, self._early = self.__dict__.get('_early'), None
        mirror = tail = None
        if early is not None and early.get('token') is tuples:
            # streamed backup: the rows of the first tuples were assembled (and their read-back started) before the last chunk of
            # beliefs was selected -- same kernel, same bytes; only the tuples that chunk added are assembled now
            u1 = early['tuples'].shape[0]
            rows, keys, mirror = early['rows'], early['keys'], early['mirror']
            if n > u1:
                tail, tail_keys = dev.backup_assemble(value_function.alpha_vector_array, self.gamma, t[u1:, 0], t[u1:, 1:], with_hash=True)
                rows, keys = torch.cat([rows, tail], dim=0), torch.cat([keys, tail_keys], dim=0)
        else:
            rows, keys = dev.backup_assemble(value_function.alpha_vector_array, self.gamma, t[:, 0], t[:, 1:], with_hash=True)
        gfirst, owner, inverse = dev.group_keys(keys, rank=rank, want_inverse=True)
        if gfirst.shape[0] == n:
            actions, hashes = t[:, 0], keys
            if mirror is not None:
                mirror = dev.mirror_finish(mirror, tail)
        elif dev.confirm_groups(rows, gfirst, inverse):
            gf = gfirst.long()
            actions, rows, hashes = t[owner.long(), 0], rows[gf], keys[gf]
        else:                                                        # 128-bit key collision (never observed): group on the bytes
            first_h, _, _, ginv = _dedup_exact_host(rows, keys.cpu().numpy())
            last_h = rank.cpu().numpy()
            order = np.lexsort((last_h, ginv))                       # within a byte group: ascending position of the last belief
            ends = np.append(np.flatnonzero(np.diff(ginv[order])), order.shape[0] - 1)
            own = np.empty(first_h.shape[0], dtype=np.int64)
            own[ginv[order[ends]]] = order[ends]
            gf = torch.as_tensor(first_h, device=dev.device)
            actions, rows, hashes = t[torch.as_tensor(own, device=dev.device), 0], rows[gf], keys[gf]
        out = ValueFunction(model, rows, actions.cpu().numpy().astype(np.int64), _trusted=True, _hashes=hashes)
        if mirror is not None and gfirst.shape[0] == n:
            out._mirror = mirror          # `numpy(staged=True)` returns this host copy (already on its way) instead of reading the rows back
        return out

    SMALL_PATH = True              # use the single-kernel backup where the sizes qualify (False: always the general pipeline)
    STREAM_FIRST_CHUNK = 1024      # rows of the first host->device chunk (small, so the score kernel starts early)
    STREAM_CHUNK = 3072            # rows of the later chunks (large, so each launch fills the 148 SMs for many waves); the plan doubles up to it
    EARLY_ROWS = True              # streamed backup: assemble + read back the rows known before the last chunk while that chunk is scored
    EARLY_MIN_TUPLES = 64

    def _chunk_plan(self, nB: int, unit: int = 1) -> list:
        """Row ranges of the streamed select: STREAM_FIRST_CHUNK rows, then doubling up to STREAM_CHUNK, all multiples of `unit`."""
        first = max(1, self.STREAM_FIRST_CHUNK // unit) * unit
        cap = max(first, int(round(self.STREAM_CHUNK / unit)) * unit)
        bounds, lo, size = [], 0, first
        while lo < nB:
            hi = min(nB, lo + size)
            if nB - hi < first:                       # a tail shorter than the first chunk joins the last one
                hi = nB
            bounds.append((lo, hi))
            lo, size = hi, min(cap, 2 * size)
        return bounds

    def _select_streamed(self, model: Model, belief_set: BeliefSet, V: torch.Tensor, want_value: bool = True, early: dict | None = None):
        """
        Select step for a belief set that still lives in (pinned) host memory: the rows are uploaded in chunks on a
        copy stream while the compute stream runs `pbvi_backup_select` on the chunks that have landed (rows are
        independent given V), so the PCIe transfer hides behind the score kernel instead of preceding it.

        Sparse rows travel packed: host threads turn slabs of rows into [bitmap | non-zero 4-double chunks]
        (`pbvi_pack_slabs_host`; started when the host-resident BeliefSet was created, `DeviceModel.start_pack`), only that
        crosses the bus and `pbvi_unpack_rows` rebuilds the dense rows in HBM, byte for byte, on a stream of its own -- beside the
        score kernel of the previous chunk.  The packers run ahead of the copies, the copies ahead of the kernels.
        `last_h2d_bytes` = bytes copied.

        `early` (a dict to fill, or None): before the LAST chunk is selected, the distinct tuples of the rows selected so far are
        grouped, their alpha rows assembled and the device->host copy of those rows started (`DeviceModel.mirror_begin`) -- the
        link is full duplex and the score kernel of the last chunk hides the copy.  The grouping synchronises on the chunks before the
        last one; the last chunk's upload and unpack are enqueued before that, its select right after.
        """
        import time
        dev = model.device
        host = belief_set._host
        nB, S = host.shape
        full = torch.empty((nB, S), dtype=torch.float64, device=dev.device)
        vstar = torch.empty((nB, dev.A, dev.O), dtype=torch.int32, device=dev.device)
        value = torch.empty((nB, dev.A), dtype=torch.float64, device=dev.device) if want_value else None
        astar = torch.empty((nB,), dtype=torch.int32, device=dev.device)
        compute = torch.cuda.current_stream(dev.device)
        trace = self.__dict__.get('stream_trace')       # tools/e2e_timeline.py: a list to fill with per-chunk events (None: no tracing)

        def select(lo, hi, landed):
            v, val, a = dev.backup_select(full[lo:hi], V, self.gamma, want_value=want_value)
            vstar[lo:hi], astar[lo:hi] = v, a
            if want_value:
                value[lo:hi] = val
            if trace is not None:
                done_ev = torch.cuda.Event(enable_timing=True)
                done_ev.record(compute)
                trace.append({'rows': (lo, hi), 'enqueued_host_s': time.perf_counter(), 'copied': landed, 'selected': done_ev})

        # the upload was started when the host-resident BeliefSet was created (`DeviceModel.start_pack`); start it now otherwise
        job = belief_set.__dict__.pop('_pack_job', None) or dev.start_pack(host)
        if job is not None and job.is_packed():
            st = job.st
            SL, region = st['SL'], st['region']
            unpack = st.get('unpack_stream')            # the rows are rebuilt beside the score kernel of the previous chunk, not before this one's
            if unpack is None:
                unpack = st['unpack_stream'] = torch.cuda.Stream(device=dev.device)
            unpack.wait_stream(compute)                 # `full` must be allocated before the unpack kernels touch it
            plan = self._chunk_plan(nB, job.ship_rows)
            for lo, hi in plan:
                ev = job.shipped(hi)                    # waits for the packers slab by slab and enqueues the copies
                s0, s1 = lo // SL, -(-hi // SL)
                with torch.cuda.stream(unpack):
                    unpack.wait_event(ev)
                    dev.unpack_rows(st['d_bm'][lo:hi], st['d_rs'][s0:s1], st['d_pk'][s0 * region:], full[lo:hi], slab_rows=SL, region_chunks=region // 4)
                    rebuilt = torch.cuda.Event(enable_timing=trace is not None)
                    rebuilt.record(unpack)
                if early is not None and len(plan) >= 3 and hi == nB:
                    self._early_rows(dev, V, vstar, astar, lo, early)
                compute.wait_event(rebuilt)
                select(lo, hi, rebuilt)
            job.close()
            self.last_h2d_bytes = job.h2d_bytes
        else:
            # rows as they are, chunk by chunk (dense belief sets; several ranks per host: the packers would compete for its memory)
            if job is not None:
                job.close()
            copy = getattr(self, '_copy_stream', None)
            if copy is None:
                copy = self._copy_stream = torch.cuda.Stream(device=dev.device)
            copy.wait_stream(compute)                   # `full` must be allocated before the copies touch it
            self.last_h2d_bytes = 0
            for lo, hi in self._chunk_plan(nB):
                with torch.cuda.stream(copy):
                    full[lo:hi].copy_(host[lo:hi], non_blocking=True)
                    self.last_h2d_bytes += (hi - lo) * S * 8
                    ev = torch.cuda.Event(enable_timing=trace is not None)
                    ev.record(copy)
                compute.wait_event(ev)
                select(lo, hi, ev)
        belief_set._device = full
        return vstar, value, astar

    def _early_rows(self, dev, V: torch.Tensor, vstar: torch.Tensor, astar: torch.Tensor, n_done: int, early: dict) -> None:
        """Distinct tuples of the first `n_done` beliefs (selected, in stream order) -> their alpha rows -> read-back started."""
        ar = torch.arange(n_done, device=dev.device)
        keys = torch.cat([astar[:n_done, None], vstar[ar, astar[:n_done].long()]], dim=1)
        first, _, _ = dev.group_keys(keys)              # synchronises: the chunks before the last one are done
        if first.shape[0] < self.EARLY_MIN_TUPLES:
            return
        t = keys[first.long()]
        rows, hk = dev.backup_assemble(V, self.gamma, t[:, 0], t[:, 1:], with_hash=True)
        early.update(tuples=t, rows=rows, keys=hk, mirror=dev.mirror_begin(rows))

    # ------------------------------------------------------------------------------------------------------------
    def compute_change(self, value_function: ValueFunction, new_value_function: ValueFunction, belief_set: BeliefSet) -> float:
        """max_b | max_v b.alpha_v - max_v' b.alpha'_v' | (reference src/pomdp.py:2141-2169)."""
        if len(belief_set) == 0:
            return 0.0
        a = self._max_values_cached(value_function, belief_set)
        b = self._max_values_cached(new_value_function, belief_set)
        return float(torch.max(torch.abs(b - a)))

    def _max_values_cached(self, vf: ValueFunction, belief_set: BeliefSet) -> torch.Tensor:
        """
        max_v b.alpha_v for every belief of the set, reusing what earlier calls already computed.  The solve loop evaluates
        the same value functions on a belief set that only grows by appended rows (`BeliefSet.union` keeps a lineage), and a
        new-points backup returns `new rows + all old rows` (`parent_uid`), so only two thin products are ever new:
        (all beliefs) x (new rows) and (new beliefs) x (all rows).  Each b.alpha is produced by the same kernel whatever the
        batch it is computed in (skipped chunks contribute exact zeros), so the result equals the full recomputation bit for bit.
        """
        dev = belief_set.model.device
        cache = self.__dict__.setdefault('_max_cache', {})
        B = belief_set.belief_array
        nB = B.shape[0]
        chain = belief_set.lineage_chain        # every union result has its own lineage; the chain says which prefixes it shares

        def known(uid):
            """Longest cached prefix of per-belief maxima of value function `uid` that is valid for the rows of this set."""
            for lineage, n in chain:
                hit = cache.get((uid, lineage))
                if hit is not None:
                    return hit[:min(n, hit.shape[0], nB)]
            return None

        key = (vf.uid, belief_set.lineage)
        done = known(vf.uid)
        if done is None and vf.parent_uid is not None and 0 < vf.n_new:
            parent = known(vf.parent_uid)
            if parent is not None and parent.shape[0] > 0:
                fresh, _ = dev.max_values(B[:parent.shape[0]], vf.alpha_vector_array[:vf.n_new])
                done = torch.maximum(parent, fresh)
        if done is None:
            done = torch.empty((0,), dtype=torch.float64, device=dev.device)
        if done.shape[0] > nB:
            done = done[:nB]
        if done.shape[0] < nB:
            rest, _ = dev.max_values(B[done.shape[0]:], vf.alpha_vector_array)
            done = torch.cat([done, rest])
        cache[key] = done
        while len(cache) > 8:                       # a few value functions are alive at any time in the solve loop
            cache.pop(next(iter(cache)))
        return done

    # ---- expansions ------------------------------------------------------------------------------------------------
    def expand_ra(self, model: Model, belief_set: BeliefSet, max_generation: int = 10) -> BeliefSet:
        """Random beliefs (reference src/pomdp.py:1527-1548); host NumPy RNG."""
        generation_count = min(len(belief_set), max_generation)
        new_beliefs = np.random.random((generation_count, model.state_count))
        new_beliefs /= np.sum(new_beliefs, axis=1)[:, None]
        return BeliefSet(model, new_beliefs)

    def _stochastic_step(self, model: Model, belief_set: BeliefSet, max_generation: int, choose_action) -> BeliefSet:
        """Shared body of SSRA / SSGA: the host draws (s, a, s', o) per selected belief in the reference's order, then
        ONE batched belief-update launch replaces the per-belief `b.update(a, o)`."""
        n_old = len(belief_set)
        to_generate = min(max_generation, n_old)
        rand_ind = np.random.choice(np.arange(n_old), to_generate, replace=False)
        src = belief_set.belief_array[torch.as_tensor(rand_ind, device=belief_set.belief_array.device)]
        host = src.cpu().numpy()
        acts, obs = np.zeros(to_generate, dtype=np.int32), np.zeros(to_generate, dtype=np.int32)
        for i in range(to_generate):
            s = int(np.random.choice(a=model.states, size=1, p=host[i])[0])
            a = int(choose_action(int(rand_ind[i])))
            s_p = model.transition(s, a)
            o = model.observe(s_p, a)
            acts[i], obs[i] = a, o
        new_rows, _ = model.device.belief_update(src, acts, obs)
        return BeliefSet(model, new_rows)

    def expand_ssra(self, model: Model, belief_set: BeliefSet, max_generation: int = 10) -> BeliefSet:
        """Stochastic Simulation with Random Action (reference src/pomdp.py:1551-1592)."""
        return self._stochastic_step(model, belief_set, max_generation, lambda belief_index: random.choice(model.actions))

    def expand_ssga(self, model: Model, belief_set: BeliefSet, value_function: ValueFunction, epsilon: float = 0.1,
                    max_generation: int = 10) -> BeliefSet:
        """Stochastic Simulation with Greedy Action (reference src/pomdp.py:1595-1648): greedy a = action of argmax_v alpha_v.b."""
        _, best = model.device.max_values(belief_set.belief_array, value_function.alpha_vector_array)
        greedy = value_function.actions[best.cpu().numpy()]

        def choose(belief_index):
            if random.random() < epsilon:
                return random.choice(model.actions)
            return greedy[belief_index]
        return self._stochastic_step(model, belief_set, max_generation, choose)

    def _successors_chunked(self, model: Model, belief_set: BeliefSet, chunk: int = 256):
        B = belief_set.belief_array
        for i0 in range(0, B.shape[0], chunk):
            succ, mass = model.device.belief_successors(B[i0:i0 + chunk])
            yield i0, succ, mass

    def expand_ssea(self, model: Model, belief_set: BeliefSet, max_generation: int = 10) -> BeliefSet:
        """
        Stochastic Simulation with Exploratory Action (reference src/pomdp.py:1651-1694): of all B*A*O successors keep the
        `max_generation` farthest (L2) from the current belief set.  Successors of impossible observations are NaN rows
        in the reference, which then crashes on every model that has one; here they are never selected.
        """
        dev = model.device
        to_generate = min(max_generation, len(belief_set))
        dists = []
        for i0, succ, mass in self._successors_chunked(model, belief_set):
            # one launch gave every successor of the chunk; only the possible ones (mass P(o|b,a) > 0, i.e. not 0/0 rows) are candidates
            flat = succ.reshape(-1, dev.S)
            possible = torch.nonzero(mass.reshape(-1) > 0)[:, 0]
            d = torch.full((flat.shape[0],), float('-inf'), dtype=torch.float64, device=dev.device)
            if possible.numel():
                cand = flat if possible.numel() == flat.shape[0] else flat[possible]
                d[possible] = dev.min_l2_distance(belief_set.belief_array, cand)
            dists.append(d.reshape(succ.shape[:3]).cpu().numpy())
        dist = np.concatenate(dists, axis=0)
        dist = np.where(np.isnan(dist), -np.inf, dist)
        pick = np.argsort(dist, axis=None)[::-1][:to_generate]
        pick = pick[np.isfinite(dist.reshape(-1)[pick])]
        b_star, a_star, o_star = np.unravel_index(pick, dist.shape)
        src = belief_set.belief_array[torch.as_tensor(b_star, device=dev.device)]
        rows, _ = dev.belief_update(src, a_star.astype(np.int32), o_star.astype(np.int32))
        return BeliefSet(model, rows)

    def expand_ger(self, model: Model, belief_set: BeliefSet, value_function: ValueFunction, max_generation: int = 10) -> BeliefSet:
        """Greedy Error Reduction (reference src/pomdp.py:1697-1765); impossible-observation successors weigh 0."""
        dev = model.device
        to_generate = min(max_generation, len(belief_set))
        r_min = model._min_reward / (1 - self.gamma)
        r_max = model._max_reward / (1 - self.gamma)
        B, V = belief_set.belief_array, value_function.alpha_vector_array
        _, best = dev.max_values(B, V)
        b_alphas = V[best.long()]
        eps_l, prob_l = [], []
        for i0, succ, mass in self._successors_chunked(model, belief_set):
            eps_l.append(dev.ger_scores(B[i0:i0 + succ.shape[0]], b_alphas[i0:i0 + succ.shape[0]], succ, r_min, r_max).cpu().numpy())
            prob_l.append(dev.observation_probabilities(B[i0:i0 + succ.shape[0]]).cpu().numpy())
        eps, bao_probs = np.concatenate(eps_l), np.concatenate(prob_l)
        res = np.einsum('bao,bao->ba', bao_probs, eps)
        b_stars, a_stars = np.unravel_index(np.argsort(res, axis=None)[::-1][:to_generate], res.shape)
        o_star = np.argmax(bao_probs[b_stars, a_stars] * eps[b_stars, a_stars], axis=1)
        src = B[torch.as_tensor(b_stars, device=dev.device)]
        rows, _ = dev.belief_update(src, a_stars.astype(np.int32), o_star.astype(np.int32))
        return BeliefSet(model, rows)

    def expand_hsvi(self, model: Model, b: Belief, value_function: ValueFunction, upper_bound_belief_value_map: BeliefValueMapping,
                    conv_term: Union[float, None] = None, max_generation: int = 10) -> BeliefSet:
        """
        HSVI exploration (reference src/pomdp.py:1768-1868): a = argmax of the upper-bound Q, o = argmax P(o|b,a) * (upper - lower),
        recursing until the gap closes or `max_generation` is reached.  Observations with P(o|b,a) = 0 are skipped
        (their successor is 0/0 in the reference).  The reference rebuilds a BeliefSet at every level of the unwinding
        recursion (:1862-1868); here the levels hand a list up and the set is stacked once -- same beliefs, same order.
        """
        return BeliefSet(model, self._expand_hsvi_list(model, b, value_function, upper_bound_belief_value_map, conv_term, max_generation))

    def _expand_hsvi_list(self, model: Model, b: Belief, value_function: ValueFunction, upper_bound_belief_value_map: BeliefValueMapping,
                          conv_term: Union[float, None], max_generation: int) -> list:
        """The recursion of the reference unrolled: one `pbvi_hsvi_level` call (one synchronisation) per level; the beliefs come out
        in the recursion's order (deepest first)."""
        dev = model.device
        ub = upper_bound_belief_value_map
        if conv_term is None:
            conv_term = self.eps
        chosen = []
        V = value_function.alpha_vector_array.contiguous()
        # the chosen successors of all levels land in one buffer, written by the level call itself (no allocation, no copy per level)
        chain = torch.empty((max(int(max_generation), 1), model.state_count), dtype=torch.float64, device=dev.device)
        level = 0
        while True:
            conv_term /= self.gamma
            idx, val, count, dot, vals, n_ub = ub._arrays()
            n_stored = len(ub.beliefs)
            ub._reserve(n_stored + 1)
            may_continue = max_generation > 1
            _, _, res, meta = dev.hsvi_level(b.values, V, self.gamma, ub.corner_values, idx, val, count, dot, vals, n_ub,
                                             ub._keys_dev, ub._vals_dev, n_stored, conv_term, may_continue, next_out=chain[level],
                                             want_successors=False)
            best_a, best_o, max_qv, best_v_diff = int(res[0]), int(res[1]), float(res[2]), float(res[3])
            next_b = b if best_o < 0 else Belief._from_device(model, chain[level])
            level += 1
            if best_v_diff < conv_term or not may_continue:
                chosen.append(next_b)
                break
            key = (int(meta[1]), int(meta[2]))
            b._row_key = key
            if int(meta[0]):
                ub._append(b, key, max_qv, on_device=True)
            elif key not in ub.belief_value_mapping:
                ub._append(b, key, max_qv, on_device=False)          # (stored arrays were full: host-side append)
            chosen.append(next_b)
            b = next_b
            max_generation -= 1
        chosen.reverse()
        return chosen

    def _trajectory(self, model: Model, b0: Belief, mdp_policy: ValueFunction, max_generation: int, eps_greedy=None) -> BeliefSet:
        """
        FSVI / FSVI_EG trajectory (reference src/pomdp.py:1871-2007).  The state walk only needs host draws, so the whole
        (a, o) sequence is drawn first -- same generators, same order as the reference -- and the belief chain is then
        advanced on the device with no host round trip.  `a_star` is the ROW index of the best MDP alpha vector at s, used
        as the action, exactly as the reference does (`xp.argmax(mdp_policy.alpha_vector_array[:, s])`).
        """
        dev = model.device
        policy_rows = torch.argmax(mdp_policy.alpha_vector_array, dim=0).cpu().numpy()
        b0_host = b0.values_host
        s = int(np.random.choice(a=model.states, size=1, p=b0_host)[0])
        steps = []
        for i in range(max_generation - 1):
            if eps_greedy is not None and random.random() < eps_greedy(i):
                a_star = int(random.choice(model.actions))
            else:
                a_star = int(policy_rows[s])
            s_p = model.transition(s, a_star)
            o = model.observe(s_p, a_star)
            reset = s_p in model.end_states
            steps.append((a_star, o, reset))
            s = s_p
            if reset:
                s = int(np.random.choice(a=model.states, size=1, p=b0_host)[0])
        if not steps:
            return BeliefSet(model, b0.values[None, :].clone())
        chain = dev.belief_trajectory(b0.values, [st[0] for st in steps], [st[1] for st in steps], [st[2] for st in steps])
        return BeliefSet(model, torch.cat([b0.values[None, :], chain], dim=0))

    def expand_fsvi(self, model: Model, b0: Belief, mdp_policy: ValueFunction, max_generation: int = 10) -> BeliefSet:
        """Forward Search Value Iteration exploration (reference src/pomdp.py:1871-1935)."""
        return self._trajectory(model, b0, mdp_policy, max_generation)

    def expand_fsvi_eg(self, model: Model, b0: Belief, mdp_policy: ValueFunction, eps_greedy=None, max_generation: int = 10) -> BeliefSet:
        """FSVI with epsilon-greedy actions (reference src/pomdp.py:1938-2007)."""
        if eps_greedy is None:
            eps_greedy = (lambda t: 0.2)
        return self._trajectory(model, b0, mdp_policy, max_generation, eps_greedy)

    def expand_perseus(self, model: Model, b: Belief, max_generation: int = 10) -> BeliefSet:
        """
        Random walk in belief space (reference src/pomdp.py:2010-2056): a uniform, o ~ P(o|b,a), b <- update(b,a,o).
        The host draws, per step and in the reference's order, the action (`np.random.choice(model.actions)`) and the ONE uniform
        that `np.random.choice(observations, p=obs_prob)` consumes; the observation itself is picked on the device with NumPy's
        rule (cumsum, normalise, searchsorted 'right'), so the whole walk is enqueued without a host round trip per step and the
        legacy RNG stream advances exactly as in the reference.
        """
        dev = model.device
        n = int(max_generation)
        if n <= 0:
            return BeliefSet(model, torch.empty((0, model.state_count), dtype=torch.float64, device=dev.device))
        acts, us = np.empty(n, dtype=np.int32), np.empty(n, dtype=np.float64)
        n_actions = len(model.actions)
        for i in range(n):
            # `np.random.choice(model.actions, size=1)` draws its index with `randint(0, len, size=1)`: the same call, without choice's
            # argument checks (half the host time of a long walk); `choice(observations, p=...)` consumes one `random_sample()`
            acts[i] = int(model.actions[np.random.randint(0, n_actions)])
            us[i] = np.random.random_sample()
        return BeliefSet(model, dev.perseus_walk(b.values, acts, us))

    def expand(self, model: Model, belief_set: BeliefSet, max_generation: int, **function_specific_parameters) -> BeliefSet:
        """Dispatcher (reference src/pomdp.py:2059-2138): the strategy is matched by substring, e.g. 'ra', 'ssra', 'expand_ssra'."""
        p = function_specific_parameters
        if self.expand_function in 'expand_ra':
            return self.expand_ra(model=model, belief_set=belief_set, max_generation=max_generation)
        elif self.expand_function in 'expand_ssra':
            return self.expand_ssra(model=model, belief_set=belief_set, max_generation=max_generation)
        elif self.expand_function in 'expand_ssga':
            args = {arg: p[arg] for arg in ['value_function', 'epsilon'] if arg in p}
            return self.expand_ssga(model=model, belief_set=belief_set, max_generation=max_generation, **args)
        elif self.expand_function in 'expand_ssea':
            return self.expand_ssea(model=model, belief_set=belief_set, max_generation=max_generation)
        elif self.expand_function in 'expand_ger':
            args = {arg: p[arg] for arg in ['value_function'] if arg in p}
            return self.expand_ger(model=model, belief_set=belief_set, max_generation=max_generation, **args)
        elif self.expand_function in 'expand_hsvi':
            args = {arg: p[arg] for arg in ['value_function', 'mdp_policy'] if arg in p}
            if not hasattr(self, '_upper_bound'):
                self._upper_bound = BeliefValueMapping(model, args['mdp_policy'])
            else:
                self._upper_bound.update()
            return self.expand_hsvi(model=model, b=belief_set.belief_at(0), value_function=args['value_function'],
                                    upper_bound_belief_value_map=self._upper_bound, max_generation=max_generation)
        elif self.expand_function in 'expand_fsvi':
            return self.expand_fsvi(model=model, b0=belief_set.belief_at(0), mdp_policy=p['mdp_policy'], max_generation=max_generation)
        elif self.expand_function in 'expand_fsvi_eg':
            return self.expand_fsvi_eg(model=model, b0=belief_set.belief_at(0), mdp_policy=p['mdp_policy'],
                                       eps_greedy=p.get('eps_greedy'), max_generation=max_generation)
        elif self.expand_function in 'expand_perseus':
            return self.expand_perseus(model=model, b=belief_set.belief_at(0), max_generation=max_generation)
        raise Exception('Not implemented')

    # ------------------------------------------------------------------------------------------------------------
    def solve(self, model: Model, expansions: int, full_backup: Union[bool, None] = None, update_passes: int = 1,
              max_belief_growth: int = 10, initial_belief: Union[BeliefSet, Belief, None] = None,
              initial_value_function: Union[ValueFunction, None] = None, prune_level: int = 1, prune_interval: int = 10,
              limit_value_function_size: int = -1, use_gpu: bool = True, history_tracking_level: int = 1,
              print_progress: bool = True, *, group=None, replicate_below: int = 64):
        """
        Expand / backup loop (reference src/pomdp.py:2172-2413), same arguments, same control flow.  Always runs on the device.
        Returns (ValueFunction, SolverHistory).

        `group` (keyword-only, not in the reference): a torch.distributed process group (or True for the default group) over
        which the loop is sharded, one process per GPU -- see `parallel.ShardedSolveState`: expansion on rank 0 + broadcast of
        the new rows, append-only row ownership, belief-sharded backups with the tuple exchange, sharded `compute_change`.  Every
        rank calls `solve` with the same arguments and gets the same value function and history counts as a single process
        (host RNG draws happen on rank 0 only).
        """
        shard = None
        if group is not None and group is not False:
            from .parallel import ShardedSolveState
            shard = ShardedSolveState(self, model, None if group is True else group, replicate_below=replicate_below)
            if shard.world == 1:
                shard = None
        if initial_belief is None:
            belief_set = BeliefSet(model, [Belief(model)])
        elif isinstance(initial_belief, BeliefSet):
            belief_set = initial_belief
        else:
            belief_set = BeliefSet(model, [Belief(model, initial_belief.values)])
        if initial_value_function is None:
            value_function = ValueFunction(model, model.expected_rewards_table.T, model.actions)
        else:
            value_function = initial_value_function
        if full_backup is None:
            full_backup = any([self.expand_function in func for func in ['expand_ra', 'expand_ssra', 'expand_ssga', 'expand_ssea', 'expand_ger']])
        if (('fsvi' in self.expand_function or 'hsvi' in self.expand_function) and
                (('mdp_policy' not in self.expand_function_params) or (self.expand_function_params['mdp_policy'] is None))):
            log('[Warning] MDP solution not provided, running value iteration on the problem to retrieve it...')
            vi_solver = VI_Solver(gamma=self.gamma, eps=self.eps)
            log('    > Starting MDP Value Iteration...')
            mdp_solution, hist = vi_solver.solve(model, use_gpu=use_gpu, print_progress=False)
            log(f'    > Value Iteration stopped or converged in {sum(hist.iteration_times):.3f}s, and after {len(hist.iteration_times)} iteration.\n')
            self.expand_function_params['mdp_policy'] = mdp_solution

        self.__dict__.pop('_max_cache', None)        # per-belief maxima cached by compute_change belong to one solve
        max_allowed_change = self.eps * (self.gamma / (1 - self.gamma))
        solver_history = SolverHistory(tracking_level=history_tracking_level, model=model, gamma=self.gamma, eps=self.eps,
                                       expand_function=self.expand_function, expand_append=full_backup,
                                       initial_value_function=value_function, initial_belief_set=belief_set)
        iteration = 0
        expand_value_function = value_function
        old_value_function = value_function
        if shard is not None:
            shard.absorb(belief_set)
        self._shard_state = shard
        try:
            if print_progress:
                from tqdm.auto import trange
                iterator = trange(expansions, desc='Expansions')
            else:
                iterator = range(expansions)
            iterator_postfix = {}
            for expansion_i in iterator:
                # 1: expand the belief set
                start_ts = _now_synced(model)
                if shard is None:
                    new_belief_set = self.expand(model=model, belief_set=belief_set, value_function=value_function,
                                                 max_generation=max_belief_growth, **self.expand_function_params)
                else:
                    new_belief_set = shard.expand(model, belief_set, value_function, max_belief_growth, **self.expand_function_params)
                belief_set = belief_set.union(new_belief_set)
                if shard is not None:
                    shard.absorb(belief_set)
                solver_history.add_expand_step(expansion_time=(_now_synced(model) - start_ts).total_seconds(), belief_set=belief_set)

                # 2: backup
                for _ in range(update_passes):
                    start_ts = _now_synced(model)
                    if shard is None:
                        value_function = self.backup(model, belief_set if full_backup else new_belief_set, value_function,
                                                     append=(not full_backup), belief_dominance_prune=False)
                    elif full_backup:
                        value_function = shard.backup_full(value_function)
                    else:
                        value_function = shard.backup_new(new_belief_set, value_function)
                    backup_time = (_now_synced(model) - start_ts).total_seconds()

                    if (iteration % prune_interval) == 0 and iteration > 0:
                        start_ts = _now_synced(model)
                        vf_len = len(value_function)
                        value_function.prune(prune_level)
                        solver_history.add_prune_step((_now_synced(model) - start_ts).total_seconds(), len(value_function) - vf_len)

                    if limit_value_function_size >= 0 and len(value_function) > limit_value_function_size:
                        if shard is None:
                            value_function, n_useful = self._limit_value_function(model, value_function, belief_set, max_belief_growth)
                        else:
                            value_function, n_useful = shard.limit_value_function(value_function, belief_set, max_belief_growth)
                        iterator_postfix['|useful|'] = n_useful

                    max_change = (self.compute_change(value_function, old_value_function, belief_set) if shard is None
                                  else shard.compute_change(value_function, old_value_function))
                    solver_history.add_backup_step(backup_time, max_change, value_function)
                    if max_change < max_allowed_change:
                        break
                    old_value_function = value_function
                    iteration += 1

                expand_max_change = (self.compute_change(expand_value_function, value_function, belief_set) if shard is None
                                     else shard.compute_change(expand_value_function, value_function))
                if expand_max_change < max_allowed_change:
                    print('Converged!')
                    break
                expand_value_function = value_function
                iterator_postfix['|V|'] = len(value_function)
                iterator_postfix['|B|'] = len(belief_set)
                if print_progress:
                    iterator.set_postfix(iterator_postfix)
        except MemoryError as e:
            print(f'Memory full: {e}')
            print('Returning value function and history as is...\n')

        start_ts = _now_synced(model)
        vf_len = len(value_function)
        value_function.prune(prune_level)
        solver_history.add_prune_step((_now_synced(model) - start_ts).total_seconds(), len(value_function) - vf_len)
        return value_function, solver_history

    def _limit_value_function(self, model: Model, value_function: ValueFunction, belief_set: BeliefSet, max_belief_growth: int,
                              return_keep: bool = False):
        """Drops `max_belief_growth` randomly chosen alpha vectors that are best at no belief (reference src/pomdp.py:2347-2367;
        sampled WITH replacement and a linearly decaying weight, like the reference)."""
        _, best = model.device.max_values(belief_set.belief_array, value_function.alpha_vector_array)
        useful = np.unique(best.cpu().numpy())
        unuseful = np.delete(np.arange(len(value_function)), useful)
        weights = np.arange(len(unuseful))[::-1] / np.sum(np.arange(len(unuseful)))
        to_delete = np.random.choice(unuseful, size=max_belief_growth, p=weights)
        keep = np.delete(np.arange(len(value_function)), to_delete)
        if return_keep:
            return keep, useful.shape[0]
        rows = value_function.alpha_vector_array[torch.as_tensor(keep, device=model.device.device)]
        return ValueFunction(model, rows, value_function.actions[keep], _trusted=True, _hashes=value_function.row_hashes[keep]), useful.shape[0]


class HSVI_Solver(PBVI_Solver):
    """Heuristic Search Value Iteration preset (reference src/pomdp.py:2416-2479): new-points backup, one pass."""

    def __init__(self, gamma: float = 0.99, eps: float = 0.001, mdp_solution: Union[ValueFunction, None] = None):
        super().__init__(gamma=gamma, eps=eps, expand_function='hsvi', mdp_policy=mdp_solution)

    def solve(self, model: Model, expansions: int, max_belief_growth: int = 10, initial_belief=None, initial_value_function=None,
              prune_level: int = 1, prune_interval: int = 10, limit_value_function_size: int = -1, use_gpu: bool = True,
              history_tracking_level: int = 1, print_progress: bool = True, **sharding):
        return super().solve(model=model, expansions=expansions, full_backup=False, update_passes=1, max_belief_growth=max_belief_growth,
                             initial_belief=initial_belief, initial_value_function=initial_value_function, prune_level=prune_level,
                             prune_interval=prune_interval, limit_value_function_size=limit_value_function_size, use_gpu=use_gpu,
                             history_tracking_level=history_tracking_level, print_progress=print_progress, **sharding)


class FSVI_Solver(PBVI_Solver):
    """Forward Search Value Iteration preset (reference src/pomdp.py:2482-2545); note the default gamma of 0.9."""

    def __init__(self, gamma: float = 0.9, eps: float = 0.001, mdp_policy: Union[ValueFunction, None] = None):
        super().__init__(gamma=gamma, eps=eps, expand_function='fsvi', mdp_policy=mdp_policy)

    def solve(self, model: Model, expansions: int, max_belief_growth: int = 10, initial_belief=None, initial_value_function=None,
              prune_level: int = 1, prune_interval: int = 10, limit_value_function_size: int = -1, use_gpu: bool = True,
              history_tracking_level: int = 1, print_progress: bool = True, **sharding):
        return PBVI_Solver.solve(self, model=model, expansions=expansions, full_backup=False, update_passes=1,
                                 max_belief_growth=max_belief_growth, initial_belief=initial_belief,
                                 initial_value_function=initial_value_function, prune_level=prune_level, prune_interval=prune_interval,
                                 limit_value_function_size=limit_value_function_size, use_gpu=use_gpu,
                                 history_tracking_level=history_tracking_level, print_progress=print_progress, **sharding)


class FSVI_EG_Solver(FSVI_Solver):
    """FSVI with epsilon-greedy exploration (reference src/pomdp.py:2548-2578)."""

    def __init__(self, gamma: float = 0.9, eps: float = 0.001, mdp_policy: Union[ValueFunction, None] = None, eps_greedy=None):
        super().__init__(gamma, eps, mdp_policy)
        self.expand_function = 'fsvi_eg'
        self.expand_function_params['eps_greedy'] = eps_greedy if eps_greedy is not None else (lambda t: 0.2)
