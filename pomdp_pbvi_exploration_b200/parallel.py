"""
Belief-sharded backup over the GPUs of one box (BASELINE.json north_star item 4).

Every output row of the backup depends on one belief row, the whole (replicated) alpha set and the (replicated) model,
so the belief set is split into contiguous row blocks, one per rank, with no collective on the data path.  The only
exchange step is after the local backup: each rank holds n_r new alpha rows (after its local byte-dedup) and the ranks
all-gather them over NCCL (NVLink 5 / NVSwitch) -- the 24-byte (action, key) records first, then only the globally first
copy of every row (`exchange_new_rows`).  The merge is the same deterministic function of the rank-ordered records on every
rank, and rank order equals belief order, so the merged value function is identical on all ranks and identical to the
single-GPU result (first position, last action).

One process per GPU (torchrun); `torch.distributed` must be initialised by the caller ("nccl" for CUDA tensors; the
host-side logic is exercised with "gloo" on CPU in tests/test_parallel_gloo.py).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

from .sets import group_by_key, unique_rows_first


def shard_bounds(n_rows: int, world_size: int, rank: int) -> tuple:
    """Contiguous block [lo, hi) of rank `rank`: ceil(n/P) rows per rank, the tail ranks may be short or empty."""
    per = -(-n_rows // world_size)
    lo = min(n_rows, rank * per)
    return lo, min(n_rows, lo + per)


def _allgather_padded(local: torch.Tensor, counts: np.ndarray, group=None) -> torch.Tensor:
    """All-gather of per-rank row blocks with different row counts (`counts`, already known everywhere): blocks are padded to
    the largest count for ONE `all_gather_into_tensor`, then compacted.  Returns the rows of all ranks in rank order."""
    world = counts.shape[0]
    n_max = int(counts.max())
    width = local.shape[1]
    if n_max == 0:
        return local[:0]
    if local.shape[0] == n_max:
        pad = local.contiguous()
    else:
        pad = torch.zeros((n_max, width), dtype=local.dtype, device=local.device)
        pad[:local.shape[0]] = local
    gathered = torch.empty((world * n_max, width), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(gathered, pad, group=group)
    if np.all(counts == n_max):
        return gathered
    keep = np.concatenate([np.arange(r * n_max, r * n_max + int(c)) for r, c in enumerate(counts)])
    return gathered[torch.as_tensor(keep, device=local.device)]


def exchange_new_rows(rows: torch.Tensor, actions: np.ndarray, hashes: np.ndarray, rows_equal, group=None):
    """
    The exchange step after a sharded backup.  Every rank holds its new alpha rows [n_r,S] (already byte-deduped locally) with
    their actions and 128-bit keys; on return every rank holds the same merged set: the rank-ordered concatenation reduced with
    the reference's dict semantics (first position, last action; src/mdp.py:668-669) -- what a single process computes over the
    whole belief set.

    Keys first, rows second: (1) the (action, key) records are all-gathered (24 bytes per row) and grouped identically on
    every rank; (2) only the globally FIRST copy of each row travels -- each rank contributes the rows it owns to one NCCL
    all-gather over NVLink, which then already is the merged set in order; (3) every rank confirms, bytewise on the device, that
    its rows which lost to an earlier rank really equal the received copy (`rows_equal`), and the ranks agree on the outcome
    with a scalar all-reduce.  Returns (rows, actions, hashes, payload_bytes).
    """
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    dev = rows.device
    n = rows.shape[0]
    counts_t = torch.empty((world,), dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(counts_t, torch.tensor([n], dtype=torch.int64, device=dev), group=group)
    counts = counts_t.cpu().numpy()
    meta = np.empty((n, 3), dtype=np.int64)
    meta[:, 0] = actions
    meta[:, 1:] = hashes
    meta_all = _allgather_padded(torch.from_numpy(meta).to(dev), counts, group).cpu().numpy()
    if meta_all.shape[0] == 0:
        return rows[:0], actions[:0], hashes[:0], 0
    first, last, inverse = group_by_key(meta_all[:, 1:])                     # identical on every rank
    offsets = np.concatenate([[0], np.cumsum(counts)])
    lo, hi = offsets[rank], offsets[rank + 1]
    owner = np.searchsorted(offsets, first, side='right') - 1                # rank that holds the first copy of each group
    owned_counts = np.bincount(owner, minlength=world)
    mine = first[owner == rank] - lo                                         # ascending local positions
    send = rows if mine.shape[0] == n else rows[torch.as_tensor(mine, device=dev)]
    merged = _allgather_padded(send, owned_counts, group)                    # == rows of `first`, in order
    # local rows that lost to an earlier rank: confirm they are byte-identical to the copy that won
    local_groups = inverse[lo:hi]
    lost = np.flatnonzero(first[local_groups] < lo)
    ok = 1
    if lost.shape[0]:
        flags = rows_equal(rows, lost.astype(np.int32), merged, local_groups[lost].astype(np.int32))
        ok = int(bool(torch.as_tensor(flags).all()))
    ok_t = torch.tensor([ok], dtype=torch.int32, device=dev)
    dist.all_reduce(ok_t, op=dist.ReduceOp.MIN, group=group)
    if int(ok_t[0]) == 0:
        raise RuntimeError('128-bit row key collision between different alpha rows of two ranks')
    payload = int(owned_counts.max()) * rows.shape[1] * 8 * world + int(counts.max()) * 24 * world
    return merged, meta_all[last, 0].copy(), meta_all[first, 1:].copy(), payload


_LAST_MAX_COUNT = [512]


def merge_blocks_host(blocks: torch.Tensor, world: int, block_rows: int, words: int):
    """
    `DeviceModel.group_record_blocks` for a gathered buffer held in a CPU tensor: (first_rows [g], last_rows [g], max_records).
    Host logic only -- used where the exchange runs over gloo on CPU (tests/test_parallel_gloo.py); on a GPU the records never
    leave the device (`pbvi_group_record_blocks`).
    """
    b = blocks.numpy().reshape(world, block_rows, words + 2).astype(np.int64)
    counts = b[:, 0, 0]
    rows = np.concatenate([r * block_rows + 1 + np.arange(min(int(counts[r]), block_rows - 1)) for r in range(world)])
    if rows.shape[0] == 0:
        z = torch.zeros((0,), dtype=torch.int64)
        return z, z, int(counts.max())
    recs = b.reshape(-1, words + 2)[rows]
    first, _, inv = unique_rows_first(recs[:, :words])
    n = recs.shape[0]
    order = np.lexsort((np.arange(n), recs[:, words + 1], inv))          # per group: ascending (last position, index)
    ends = np.append(np.flatnonzero(np.diff(inv[order])), n - 1)
    last = np.empty(first.shape[0], dtype=np.int64)
    last[inv[order[ends]]] = order[ends]
    return torch.from_numpy(rows[first]), torch.from_numpy(rows[last]), int(counts.max())


def exchange_tuples(tuples, first, last, capacity: int, device, group=None, merge_fn=None):
    """
    The exchange step of the sharded backup in its compact form.  An alpha row of the backup is a deterministic function of
    its generating tuple (a*, v*[a*, :]) and of the replicated (model, old value function), so the ranks all-gather the
    distinct TUPLES of their shards (4*(1+O) bytes each instead of 8*S per row) together with the positions of the first / last
    belief that chose them, and every rank then assembles the merged set itself (`PBVI_Solver.rows_from_tuples`) -- bitwise the
    same rows on every rank, in the order a single process would produce.  One collective: block r of the gathered int32 buffer
    is [header (u_r); u_r records (tuple, first, last)], padded to a common number of rows.  Positions travel globalised as
    rank * capacity + local position (`capacity` = the largest shard size, the same on every rank): that is the position in the
    whole belief set when the shards are the contiguous blocks of `shard_bounds`, and order-preserving in any case.  The
    records stay on `device`; the merge is `merge_fn` (the device's `group_record_blocks`; host twin for CPU tensors), which
    also returns the largest record count of any rank.
    Returns (tuples [U, 1+O], first [U], last [U]) of the whole belief set as int32 tensors on `device`.
    """
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    device = torch.device(device)
    tuples = torch.as_tensor(tuples).to(device=device, dtype=torch.int32)
    first = torch.as_tensor(first).to(device=device, dtype=torch.int32)
    last = torch.as_tensor(last).to(device=device, dtype=torch.int32)
    if merge_fn is None:
        assert device.type == 'cpu', 'pass the device merge (DeviceModel.group_record_blocks) for records in device memory'
        merge_fn = merge_blocks_host
    u, w = tuples.shape
    assert u <= capacity, (u, capacity)
    assert capacity * world < 2 ** 31
    offset = rank * capacity
    # Most beliefs share their tuple with others, so the blocks are first sized by a guess (twice the largest count seen in the
    # previous exchange); the header carries the true count, and if any rank overflowed -- every rank sees that in the gathered
    # headers, so the decision is consistent -- the exchange is repeated once at full capacity.
    guess = min(capacity, max(256, 2 * _LAST_MAX_COUNT[0]))
    while True:
        k = min(u, guess)
        buf = torch.empty((guess + 1, w + 2), dtype=torch.int32, device=device)
        buf[0].fill_(u)
        buf[1:k + 1, :w] = tuples[:k]
        buf[1:k + 1, w] = first[:k] + offset
        buf[1:k + 1, w + 1] = last[:k] + offset
        gathered = torch.empty((world * (guess + 1), w + 2), dtype=torch.int32, device=device)
        dist.all_gather_into_tensor(gathered, buf, group=group)
        first_rows, last_rows, max_count = merge_fn(gathered, world, guess + 1, w)
        _LAST_MAX_COUNT[0] = max_count
        if max_count <= guess:
            break
        guess = capacity
    fr = first_rows.long()
    return gathered[fr, :w], gathered[fr, w], gathered[last_rows.long(), w + 1]


class _PhaseTimer:
    def __init__(self, device):
        import time
        self.time, self.device, self.phases = time, device, {}
        torch.cuda.synchronize(device)
        self.t = time.perf_counter()

    def mark(self, name):
        torch.cuda.synchronize(self.device)
        now = self.time.perf_counter()
        self.phases[name] = (now - self.t) * 1e3
        self.t = now


class ShardedBackup:
    """
    `PBVI_Solver.backup` over a belief set sharded across the ranks of `group`.

        sb = ShardedBackup(solver, model)
        lo, hi = sb.bounds(n_beliefs)                       # this rank's rows
        vf = sb.backup(BeliefSet(model, B[lo:hi]), value_function, append=False)   # same ValueFunction on every rank
    """

    def __init__(self, solver, model, group=None, exchange: str = 'tuples'):
        assert exchange in ('tuples', 'rows')
        self.solver = solver
        self.model = model
        self.group = group
        self.exchange = exchange        # 'tuples': all-gather the generating tuples, assemble everywhere; 'rows': all-gather the rows
        self._cap = (None, None)
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.last_payload_bytes = 0
        self.trace = False              # set True to record per-phase wall times (device-synchronised) in `last_phases`
        self.last_phases = {}

    def bounds(self, n_rows: int) -> tuple:
        self._fixed_cap = -(-n_rows // self.world)
        return shard_bounds(n_rows, self.world, self.rank)

    def backup(self, local_belief_set, value_function, append: bool = False, belief_dominance_prune: bool = False):
        from .value_function import ValueFunction
        t = _PhaseTimer(self.model.device.device) if self.trace else None
        if self.exchange == 'tuples':
            dev = self.model.device
            tuples, first, last = self.solver.select_tuples_device(self.model, local_belief_set, value_function, belief_dominance_prune)
            if t: t.mark('local select')
            n_local = len(local_belief_set)
            g_tuples, _, g_last = exchange_tuples(tuples, first, last, self._capacity(n_local), dev.device, self.group,
                                                  merge_fn=dev.group_record_blocks)
            self.last_payload_bytes = (min(self._cap[1], max(256, 2 * _LAST_MAX_COUNT[0])) + 1) * (tuples.shape[1] + 2) * 4 * self.world
            if t: t.mark('exchange')
            merged = self.solver.rows_from_tuples(self.model, value_function, g_tuples, g_last)
            if t: t.mark('assemble + dedup')
            if append:
                n_new = len(merged)
                merged.extend(value_function)
                merged.parent_uid, merged.n_new = value_function.uid, n_new
            if t:
                t.mark('extend')
                self.last_phases = t.phases
            return merged
        local = self.solver.backup(self.model, local_belief_set, value_function, append=False,
                                   belief_dominance_prune=belief_dominance_prune)
        if t: t.mark('local backup')
        rows, actions, hashes, self.last_payload_bytes = exchange_new_rows(local.alpha_vector_array, local.actions, local.row_hashes,
                                                                          self.model.device.rows_equal, self.group)
        if t: t.mark('exchange')
        merged = ValueFunction(self.model, rows, actions, _trusted=True, _hashes=hashes)
        if append:
            merged.extend(value_function)
        if t:
            t.mark('extend')
            self.last_phases = t.phases
        return merged

    def set_capacity(self, max_shard_rows: int) -> None:
        """Declares the largest shard size of any rank (the same value on every rank), which saves the per-call agreement."""
        self._fixed_cap = int(max_shard_rows)

    def _capacity(self, n_local: int) -> int:
        """Largest shard size over the ranks = upper bound on the records a rank contributes.  Known without communication after
        `bounds(n_rows)` / `set_capacity`; otherwise agreed with one scalar all-reduce per call (every rank takes part)."""
        fixed = getattr(self, '_fixed_cap', None)
        if fixed is not None:
            assert n_local <= fixed, f'shard of {n_local} rows exceeds the declared capacity {fixed}'
            self._cap = (n_local, fixed)
            return fixed
        c = torch.tensor([n_local], dtype=torch.int64, device=self.model.device.device)
        dist.all_reduce(c, op=dist.ReduceOp.MAX, group=self.group)
        self._cap = (n_local, int(c[0]))
        return self._cap[1]

    def compute_change(self, value_function, new_value_function, local_belief_set) -> float:
        """max over all shards of the local change: one scalar all-reduce(max)."""
        local = self.solver.compute_change(value_function, new_value_function, local_belief_set)
        t = torch.tensor([local], dtype=torch.float64, device=self.model.device.device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        return float(t[0])
