"""
Belief-sharded backup over the GPUs of one box (BASELINE.json north_star item 4).

Every output row of the backup depends on one belief row, the whole (replicated) alpha set and the (replicated) model,
so the belief set is split into contiguous row blocks, one per rank, with no collective on the data path.  The only
exchange step is after the local backup: each rank holds n_r new alpha rows (after its local byte-dedup) and the ranks
all-gather them -- counts first, then rows padded to the largest count -- over NCCL (NVLink 5 / NVSwitch).  Every rank
then runs the same deterministic merge in rank order, which equals belief order, so the merged value function is
identical on all ranks and identical to the single-GPU result (first position, last action).

One process per GPU (torchrun); `torch.distributed` must be initialised by the caller ("nccl" for CUDA tensors; the
host-side logic is exercised with "gloo" on CPU in tests/test_parallel_gloo.py).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

from .sets import group_by_key


def shard_bounds(n_rows: int, world_size: int, rank: int) -> tuple:
    """Contiguous block [lo, hi) of rank `rank`: ceil(n/P) rows per rank, the tail ranks may be short or empty."""
    per = -(-n_rows // world_size)
    lo = min(n_rows, rank * per)
    return lo, min(n_rows, lo + per)


def allgather_rows(rows: torch.Tensor, actions: np.ndarray, hashes: np.ndarray, group=None):
    """
    Variable-count all-gather of (rows [n_r,S] float64, actions [n_r], hashes [n_r,2]) in rank order.
    Returns (rows [sum n_r, S], actions, hashes, counts).  Two collectives: counts (P int64) and the padded payload.
    """
    world = dist.get_world_size(group)
    dev = rows.device
    S = rows.shape[1]
    n_local = torch.tensor([rows.shape[0]], dtype=torch.int64, device=dev)
    counts_t = torch.empty((world,), dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(counts_t, n_local, group=group)
    counts = counts_t.cpu().numpy()
    n_max = int(counts.max())
    if n_max == 0:
        return rows[:0], actions[:0], hashes[:0], counts
    # payload per row: S doubles of alpha + 1 (action) + 2 (hash halves), all moved as raw 8-byte words
    pad = torch.zeros((n_max, S + 3), dtype=torch.float64, device=dev)
    n = rows.shape[0]
    if n:
        pad[:n, :S] = rows
        meta = np.empty((n, 3), dtype=np.int64)
        meta[:, 0] = actions
        meta[:, 1:] = hashes
        pad[:n, S:] = torch.from_numpy(meta).view(torch.float64).to(dev)
    gathered = torch.empty((world * n_max, S + 3), dtype=torch.float64, device=dev)
    dist.all_gather_into_tensor(gathered, pad, group=group)
    keep = torch.cat([torch.arange(r * n_max, r * n_max + int(c), device=dev) for r, c in enumerate(counts)])
    gathered = gathered[keep]
    meta_all = gathered[:, S:].contiguous().view(torch.int64).cpu().numpy()
    return gathered[:, :S].contiguous(), meta_all[:, 0].copy(), meta_all[:, 1:].copy(), counts


def merge_gathered(rows: torch.Tensor, actions: np.ndarray, hashes: np.ndarray, rows_equal):
    """
    Dict-insertion merge of the rank-ordered rows: first position, last action (reference src/mdp.py:668-669).
    `rows_equal(rows, ia, rows, ib) -> flags` confirms every 128-bit key match bytewise (DeviceModel.rows_equal).
    """
    n = rows.shape[0]
    first, last, inverse = group_by_key(hashes)
    if first.shape[0] != n:
        dup = np.flatnonzero(first[inverse] != np.arange(n))
        flags = rows_equal(rows, first[inverse[dup]].astype(np.int32), rows, dup.astype(np.int32))
        if not bool(torch.as_tensor(flags).all()):
            raise RuntimeError('128-bit row key collision between different alpha rows')
        rows = rows[torch.as_tensor(first, device=rows.device)]
    return rows, actions[last], hashes[first]


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

    def __init__(self, solver, model, group=None):
        self.solver = solver
        self.model = model
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.last_payload_bytes = 0
        self.trace = False              # set True to record per-phase wall times (device-synchronised) in `last_phases`
        self.last_phases = {}

    def bounds(self, n_rows: int) -> tuple:
        return shard_bounds(n_rows, self.world, self.rank)

    def backup(self, local_belief_set, value_function, append: bool = False, belief_dominance_prune: bool = False):
        from .value_function import ValueFunction
        t = _PhaseTimer(self.model.device.device) if self.trace else None
        local = self.solver.backup(self.model, local_belief_set, value_function, append=False,
                                   belief_dominance_prune=belief_dominance_prune)
        if t: t.mark('local backup')
        rows, actions, hashes, counts = allgather_rows(local.alpha_vector_array, local.actions, local.row_hashes, self.group)
        if t: t.mark('all-gather')
        self.last_payload_bytes = int(counts.max()) * (rows.shape[1] + 3) * 8 * self.world
        rows, actions, hashes = merge_gathered(rows, actions, hashes, self.model.device.rows_equal)
        if t: t.mark('merge')
        merged = ValueFunction(self.model, rows, actions, _trusted=True, _hashes=hashes)
        if append:
            merged.extend(value_function)
        if t:
            t.mark('extend')
            self.last_phases = t.phases
        return merged

    def compute_change(self, value_function, new_value_function, local_belief_set) -> float:
        """max over all shards of the local change: one scalar all-reduce(max)."""
        local = self.solver.compute_change(value_function, new_value_function, local_belief_set)
        t = torch.tensor([local], dtype=torch.float64, device=self.model.device.device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        return float(t[0])
