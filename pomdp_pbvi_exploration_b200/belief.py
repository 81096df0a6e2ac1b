"""
Belief / BeliefSet with the reference's interface (src/pomdp.py:311-783) on device-resident storage.

`Belief.values` is a CUDA float64 tensor [S]; `Belief.update(a, o)` runs the engine's belief-update kernel, whose result
is bit-identical to the reference's bincount + np.sum normalisation (so byte-identity of beliefs, which `BeliefSet.union`
and the solver rely on, means the same thing in both engines).  `BeliefSet.belief_array` is a CUDA tensor [N,S].
"""
from __future__ import annotations

import itertools
from typing import Union

import numpy as np
import torch

from .model import Model
from .sets import dedup_rows


_LINEAGE = itertools.count(1)


def _to_device(model: Model, values) -> torch.Tensor:
    dev = model.device.device
    if isinstance(values, torch.Tensor):
        return values.to(device=dev, dtype=torch.float64).contiguous()
    return torch.as_tensor(np.ascontiguousarray(values, dtype=np.float64)).to(dev)


class Belief:
    """
    A probability distribution over the states of `model` (reference src/pomdp.py:311-486).

    Parameters
    ----------
    model : Model
    values : np.ndarray | torch.Tensor, optional
        Defaults to the model's start probabilities.  Must sum to 1 (rounded to 3 decimals), as in the reference.
    """

    def __new__(cls, *args, **kwargs):
        instance = super().__new__(cls)
        instance._bytes_repr = None
        instance._successors = {}
        instance._host = None
        return instance

    def __init__(self, model: Model, values=None):
        assert model is not None
        self.model = model
        if values is not None:
            assert values.shape[0] == model.state_count, "Belief must contain be of dimension |S|"
            self._values = _to_device(model, values)
            prob_sum = float(self._values.sum())
            rounded_sum = round(prob_sum, 3)
            assert rounded_sum == 1.0, f"States probabilities in belief must sum to 1 (found: {prob_sum}; rounded {rounded_sum})"
        else:
            self._values = model.start_belief_device

    @classmethod
    def _from_device(cls, model: Model, values: torch.Tensor) -> 'Belief':
        """No sum check, like the reference's `update` (which builds the successor through __new__, src/pomdp.py:414-416)."""
        b = cls.__new__(cls)
        b.model = model
        b._values = values
        return b

    @property
    def values(self) -> torch.Tensor:
        return self._values

    @property
    def values_host(self) -> np.ndarray:
        if self._host is None:
            self._host = self._values.cpu().numpy()
        return self._host

    @property
    def bytes_repr(self) -> bytes:
        if self._bytes_repr is None:
            self._bytes_repr = self.values_host.tobytes()
        return self._bytes_repr

    def __eq__(self, other: object) -> bool:
        return self.bytes_repr == other.bytes_repr

    def __hash__(self):
        return hash(self.bytes_repr)

    def update(self, a: int, o: int) -> 'Belief':
        """b'(s') proportional to sum_s RTO[s,a,o,.] b(s) (reference src/pomdp.py:382-421); successors are memoised."""
        succ_id = f'{a}_{o}'
        succ = self._successors.get(succ_id)
        if succ is not None:
            return succ
        out, _ = self.model.device.belief_update(self._values[None, :], [int(a)], [int(o)])
        new_belief = Belief._from_device(self.model, out[0])
        self._successors[succ_id] = new_belief
        return new_belief

    def generate_successors(self) -> list:
        """All (a,o) successors in action-major order (reference src/pomdp.py:424-438), one kernel launch pair."""
        succ, _ = self.model.device.belief_successors(self._values[None, :])
        out = []
        for a in self.model.actions:
            for o in self.model.observations:
                key = f'{a}_{o}'
                if key not in self._successors:
                    self._successors[key] = Belief._from_device(self.model, succ[0, a, o])
                out.append(self._successors[key])
        return out

    def random_state(self) -> int:
        """A state drawn from the belief with the host NumPy RNG, like the reference's CPU path (src/pomdp.py:441-452)."""
        return int(np.random.choice(a=self.model.states, size=1, p=self.values_host)[0])


class BeliefSet:
    """
    An ordered set of beliefs (reference src/pomdp.py:489-783).

    Parameters
    ----------
    model : Model
    beliefs : list[Belief] | np.ndarray | torch.Tensor
    """

    def __init__(self, model: Model, beliefs: Union[list, np.ndarray, torch.Tensor], *, _hashes: np.ndarray | None = None):
        self.model = model
        self.is_on_gpu = True
        self._belief_list = None
        self._hashes = _hashes
        self._device = None
        self._host = None
        # identity of this (immutable) row sequence for per-row result caches; `union` links the result to its parents through
        # `lineage_chain` = [(lineage, n)]: "my first n rows are the first n rows of the set with that lineage"
        self.lineage = next(_LINEAGE)
        self._chain = None
        S = model.state_count
        if isinstance(beliefs, list):
            assert all(b.values.shape[0] == S for b in beliefs), f"Beliefs in belief list provided dont all have shape ({S},)"
            self._belief_list = beliefs
            self._device = torch.stack([b.values for b in beliefs]) if len(beliefs) else \
                torch.empty((0, S), dtype=torch.float64, device=model.device.device)
        else:
            assert beliefs.shape[1] == S, f"Belief array provided doesnt have the right shape (expected (-,{S}), received {tuple(beliefs.shape)})"
            if isinstance(beliefs, torch.Tensor) and beliefs.device.type == 'cpu':
                # host tensor: kept on the host and uploaded on first device use, or streamed chunk by chunk behind the score
                # kernel by PBVI_Solver.backup (pin it -- `tensor.pin_memory()` -- for the copies to overlap the compute)
                self._host = beliefs.to(torch.float64).contiguous()
                if self._host.is_pinned() and self._host.shape[0] >= 2048:
                    # large pinned host set: the packer threads start now, so the host half of the sparse-row upload overlaps
                    # whatever the caller does before the backup (PBVI_Solver._select_streamed consumes the job)
                    self._pack_job = model.device.start_pack(self._host)
            else:
                self._device = _to_device(model, beliefs)
            if not isinstance(beliefs, torch.Tensor):
                # the reference builds Belief objects here, which asserts every row sums to 1 (src/pomdp.py:531-533, 345-347)
                sums = np.round(np.asarray(beliefs, dtype=np.float64).sum(axis=1), 3)
                bad = np.flatnonzero(~(sums == 1.0))
                assert bad.size == 0, f"States probabilities in belief must sum to 1 (found: {float(np.asarray(beliefs)[bad[0]].sum())})"

    LINEAGE_DEPTH = 8

    @property
    def lineage_chain(self) -> list:
        """[(lineage, n)], own identity first: the first n rows of this set equal the first n rows of the set `lineage`."""
        return [(self.lineage, len(self))] + (self._chain or [])

    def _inherit(self, parent: 'BeliefSet', n_prefix: int) -> None:
        self._chain = [(l, min(n, n_prefix)) for l, n in parent.lineage_chain][:self.LINEAGE_DEPTH]

    @property
    def belief_array(self) -> torch.Tensor:
        """[N,S] CUDA float64 tensor."""
        if self._device is None:
            job = self.__dict__.pop('_pack_job', None)
            if job is not None:
                job.close()                         # plain upload: the packed form is not used (the staging buffers become free)
            self._device = _to_device(self.model, self._host)
        return self._device

    @property
    def belief_list(self) -> list:
        if self._belief_list is None:
            self._belief_list = [Belief._from_device(self.model, row) for row in self.belief_array]
        return self._belief_list

    def belief_at(self, i: int) -> Belief:
        """The i-th belief without materialising `belief_list` (the solve loop only ever needs `belief_list[0]`)."""
        if self._belief_list is not None:
            return self._belief_list[i]
        return Belief._from_device(self.model, self.belief_array[i])

    @property
    def row_hashes(self) -> np.ndarray:
        if self._hashes is None:
            self._hashes = self.model.device.row_hash(self.belief_array).cpu().numpy() if len(self) else np.zeros((0, 2), dtype=np.int64)
        return self._hashes

    def __len__(self) -> int:
        return int((self._device if self._device is not None else self._host).shape[0])

    def numpy(self) -> np.ndarray:
        return self.belief_array.cpu().numpy()

    def generate_all_successors(self) -> 'BeliefSet':
        succ, _ = self.model.device.belief_successors(self.belief_array)
        return BeliefSet(self.model, succ.reshape(-1, self.model.state_count))

    def _key_index(self):
        """
        {128-bit row key: row index} of this set, or None when the set holds duplicate rows.  The dict and the row store are
        shared along a union() chain (append-only): a set may extend them only while it is the longest set using them.
        """
        idx = self.__dict__.get('_keys')
        if idx is not None and len(idx) == len(self):
            return idx
        h = self.row_hashes
        idx = {}
        for i, k in enumerate(map(tuple, h.tolist())):
            if k in idx:
                return None
            idx[k] = i
        self._keys = idx
        return idx

    def union(self, other_belief_set: 'BeliefSet') -> 'BeliefSet':
        """Own unique beliefs in order, then the unseen beliefs of the other set (reference src/pomdp.py:585-606)."""
        fast = self._union_append(other_belief_set)
        if fast is not None:
            return fast
        return self._union_general(other_belief_set)

    def _union_append(self, other: 'BeliefSet'):
        """
        The solve loop's case -- a large duplicate-free set absorbs a small one -- in O(new rows): key lookups against the
        shared index, bytewise confirmation of every hit on the device, and an append into a capacity-doubling row store
        that the resulting set shares with this one (rows are never mutated, so the prefix stays valid for `self`).
        """
        n_self, n_other = len(self), len(other)
        if n_self == 0 or n_other == 0 or n_other > n_self:
            return None
        index = self._key_index()
        if index is None:
            return None
        dev = self.model.device
        other_rows = other.belief_array
        other_keys = list(map(tuple, other.row_hashes.tolist()))
        hit_other, hit_self, fresh = [], [], []
        seen = {}
        for j, k in enumerate(other_keys):
            if k in index:
                hit_other.append(j); hit_self.append(index[k])
            elif k in seen:
                hit_other.append(j); hit_self.append(-1 - seen[k])       # duplicate inside `other`: compare with its first copy
            else:
                seen[k] = j
                fresh.append(j)
        if hit_other:
            own = [(jo, js) for jo, js in zip(hit_other, hit_self) if js >= 0]
            twin = [(jo, -1 - js) for jo, js in zip(hit_other, hit_self) if js < 0]
            ok = True
            if own:
                ok &= bool(dev.rows_equal(other_rows, [p[0] for p in own], self.belief_array, [p[1] for p in own]).all())
            if twin:
                ok &= bool(dev.rows_equal(other_rows, [p[0] for p in twin], other_rows, [p[1] for p in twin]).all())
            if not ok:
                return None                                                # 128-bit key collision: take the exact general path
        k_new = len(fresh)
        store = self.__dict__.get('_store')
        if store is None or self.__dict__.get('_store_used', [0])[0] != n_self or store.shape[0] < n_self + k_new:
            cap = max(2 * (n_self + k_new), 256)
            store = torch.empty((cap, self.model.state_count), dtype=torch.float64, device=dev.device)
            store[:n_self] = self.belief_array
            used = [n_self]
        else:
            used = self._store_used
        if k_new:
            store[n_self:n_self + k_new] = other_rows[torch.as_tensor(fresh, device=dev.device)] if k_new != n_other else other_rows
        used[0] = n_self + k_new
        hashes = np.concatenate([self.row_hashes, other.row_hashes[fresh]], axis=0) if k_new else self.row_hashes
        out = BeliefSet(self.model, store[:n_self + k_new], _hashes=hashes)
        for j in fresh:
            index[other_keys[j]] = len(index)             # == its row position: the index held exactly n_self keys on entry
        out._keys, out._store, out._store_used = index, store, used
        out._inherit(self, n_self)
        return out

    def _union_general(self, other_belief_set: 'BeliefSet') -> 'BeliefSet':
        rows = torch.cat([self.belief_array, other_belief_set.belief_array], dim=0)
        hashes = np.concatenate([self.row_hashes, other_belief_set.row_hashes], axis=0)
        first, _, _, _ = dedup_rows(self.model.device, rows, hashes)
        if first.shape[0] != rows.shape[0]:
            rows = rows[torch.as_tensor(first, device=rows.device)]
            hashes = hashes[first]
        out = BeliefSet(self.model, rows, _hashes=hashes)
        n_self = len(self)
        if first.shape[0] >= n_self and np.array_equal(first[:n_self], np.arange(n_self)):
            out._inherit(self, n_self)          # own rows survive as a prefix: per-row results cached for `self` stay valid for it
        return out

    def to_gpu(self) -> 'BeliefSet':
        return self

    def to_cpu(self) -> 'BeliefSet':
        return self
