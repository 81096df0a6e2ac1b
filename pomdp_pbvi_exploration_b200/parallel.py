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


class NativeComm:
    """
    A process group backed by the LIBRARY'S OWN NCCL binding (`pbvi_comm_*` of include/pbvi_b200.h) instead of torch.distributed: the
    collectives a host in any language would call.  Pass it wherever a `group` is accepted (`ShardedBackup`, `PBVI_Solver.solve`).
    The 128-byte id comes from `NativeComm.unique_id()` on rank 0 and reaches the other ranks by any host-side channel.
    """

    def __init__(self, model, rank: int, world: int, unique_id: bytes):
        import ctypes
        from . import _native
        self._lib = _native.load_library()
        self._check = _native._check
        self.device = model.device.device
        self._rank, self._world = int(rank), int(world)
        assert len(unique_id) == 128
        h = ctypes.c_void_p()
        buf = (ctypes.c_ubyte * 128).from_buffer_copy(unique_id)
        self._check(self._lib.pbvi_comm_init(model.device._h, buf, self._rank, self._world, ctypes.byref(h)))
        self._h = h

    @staticmethod
    def unique_id() -> bytes:
        import ctypes
        from . import _native
        buf = (ctypes.c_ubyte * 128)()
        _native._check(_native.load_library().pbvi_comm_unique_id(buf))
        return bytes(buf)

    def close(self) -> None:
        if getattr(self, '_h', None) is not None:
            self._lib.pbvi_comm_destroy(self._h)
            self._h = None

    def _stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    def all_gather(self, out: torch.Tensor, local: torch.Tensor) -> None:
        assert out.is_cuda and local.is_cuda and out.is_contiguous() and out.dtype == local.dtype
        local = local.contiguous()
        if local.dtype == torch.float64:
            self._check(self._lib.pbvi_allgather_rows(self._h, local.data_ptr(), 1, local.numel(), out.data_ptr(), self._stream()))
        else:
            words = local.numel() * local.element_size() // 4
            assert local.numel() * local.element_size() % 4 == 0
            self._check(self._lib.pbvi_allgather_tuples(self._h, local.data_ptr(), 1, words, out.data_ptr(), self._stream()))

    def all_reduce(self, t: torch.Tensor, op) -> None:
        """MAX / MIN of a small tensor of any numeric dtype, through the library's double-precision max."""
        assert op in (dist.ReduceOp.MAX, dist.ReduceOp.MIN)
        d = t.to(torch.float64)
        if op == dist.ReduceOp.MIN:
            d = -d
        d = d.contiguous()
        self._check(self._lib.pbvi_allreduce_max(self._h, d.data_ptr(), d.numel(), self._stream()))
        t.copy_((-d if op == dist.ReduceOp.MIN else d).to(t.dtype))

    def broadcast(self, t: torch.Tensor, src: int) -> None:
        assert t.is_cuda and t.is_contiguous()
        if t.dtype == torch.float64:
            self._check(self._lib.pbvi_broadcast_rows(self._h, t.data_ptr(), t.numel(), int(src), self._stream()))
        else:                                       # counts / indices: exact in float64 below 2^53
            d = t.to(torch.float64).contiguous()
            self._check(self._lib.pbvi_broadcast_rows(self._h, d.data_ptr(), d.numel(), int(src), self._stream()))
            t.copy_(d.to(t.dtype))


def group_size(group=None) -> int:
    return group._world if isinstance(group, NativeComm) else dist.get_world_size(group)


def group_rank(group=None) -> int:
    return group._rank if isinstance(group, NativeComm) else dist.get_rank(group)


def _host_backend(group=None) -> bool:
    """True when the group's collectives run on CPU tensors (gloo): device tensors are staged through the host.  Used by the
    tests that run two ranks of the real engine on ONE GPU (NCCL refuses two ranks per device)."""
    return dist.get_backend(group) == 'gloo'


def all_gather_rows(out: torch.Tensor, local: torch.Tensor, group=None) -> None:
    """`all_gather_into_tensor` that also works for CUDA tensors over a gloo group (staged through the host) and over a NativeComm."""
    if isinstance(group, NativeComm):
        group.all_gather(out, local)
    elif local.is_cuda and _host_backend(group):
        tmp = torch.empty(out.shape, dtype=out.dtype)
        dist.all_gather_into_tensor(tmp, local.cpu(), group=group)
        out.copy_(tmp)
    else:
        dist.all_gather_into_tensor(out, local, group=group)


def all_reduce_(t: torch.Tensor, op, group=None) -> torch.Tensor:
    if isinstance(group, NativeComm):
        group.all_reduce(t, op)
    elif t.is_cuda and _host_backend(group):
        tmp = t.cpu()
        dist.all_reduce(tmp, op=op, group=group)
        t.copy_(tmp)
    else:
        dist.all_reduce(t, op=op, group=group)
    return t


def broadcast_(t: torch.Tensor, src: int, group=None) -> torch.Tensor:
    """Broadcast from GROUP rank `src`."""
    if isinstance(group, NativeComm):
        group.broadcast(t, src)
        return t
    gsrc = dist.get_global_rank(group, src) if group is not None else src
    if t.is_cuda and _host_backend(group):
        tmp = t.cpu()
        dist.broadcast(tmp, gsrc, group=group)
        t.copy_(tmp)
    else:
        dist.broadcast(t, gsrc, group=group)
    return t


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
    all_gather_rows(gathered, pad, group)
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
    world, rank = group_size(group), group_rank(group)
    dev = rows.device
    n = rows.shape[0]
    counts_t = torch.empty((world,), dtype=torch.int64, device=dev)
    all_gather_rows(counts_t, torch.tensor([n], dtype=torch.int64, device=dev), group)
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
    all_reduce_(ok_t, dist.ReduceOp.MIN, group)
    if int(ok_t[0]) == 0:
        raise RuntimeError('128-bit row key collision between different alpha rows of two ranks')
    payload = int(owned_counts.max()) * rows.shape[1] * 8 * world + int(counts.max()) * 24 * world
    return merged, meta_all[last, 0].copy(), meta_all[first, 1:].copy(), payload


def merge_blocks_host(blocks: torch.Tensor, world: int, block_rows: int, words: int):
    """
    `DeviceModel.group_record_blocks` for a gathered buffer held in a CPU tensor: (first_rows [g], last_rows [g], max_records).
    Records are grouped by key; a group's first record is the one with the smallest FIRST POSITION word, its last record the one
    with the largest (last position, row), and the groups are ordered by their first position -- the dict over the whole belief
    set, whatever the assignment of beliefs to ranks.  Host logic only -- used where the exchange runs over gloo on CPU
    (tests/test_parallel_gloo.py); on a GPU the records never leave the device (`pbvi_group_record_blocks`).
    """
    b = blocks.numpy().reshape(world, block_rows, words + 2).astype(np.int64)
    counts = b[:, 0, 0]
    rows = np.concatenate([r * block_rows + 1 + np.arange(min(int(counts[r]), block_rows - 1)) for r in range(world)])
    if rows.shape[0] == 0:
        z = torch.zeros((0,), dtype=torch.int64)
        return z, z, int(counts.max())
    recs = b.reshape(-1, words + 2)[rows]
    _, _, inv = unique_rows_first(recs[:, :words])
    n = recs.shape[0]
    n_groups = int(inv.max()) + 1
    order = np.lexsort((np.arange(n), recs[:, words], inv))              # per group: ascending (first position, row)
    starts = np.append(0, np.flatnonzero(np.diff(inv[order])) + 1)
    first = np.empty(n_groups, dtype=np.int64)
    first[inv[order[starts]]] = order[starts]
    order = np.lexsort((np.arange(n), recs[:, words + 1], inv))          # per group: ascending (last position, row)
    ends = np.append(np.flatnonzero(np.diff(inv[order])), n - 1)
    last = np.empty(n_groups, dtype=np.int64)
    last[inv[order[ends]]] = order[ends]
    by_pos = np.argsort(recs[first, words], kind='stable')               # groups in order of their first position
    return torch.from_numpy(rows[first[by_pos]]), torch.from_numpy(rows[last[by_pos]]), int(counts.max())


def exchange_tuples(tuples, first, last, capacity: int, device, group=None, merge_fn=None, positions=None, guess_state=None):
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
    also returns the largest record count of any rank.  `positions` (int32 [n_local], optional) gives the position of every
    local belief in the whole belief set explicitly -- the sharded solve's append-only ownership interleaves the ranks' rows --
    and replaces the rank * capacity + local rule; the merge orders by position either way.  `guess_state` is the caller's
    one-element list holding the largest record count of its previous exchange (per `ShardedBackup`: every rank of a group
    takes part in the same sequence of exchanges, so the block size guess is the same everywhere).
    Returns (tuples [U, 1+O], first [U], last [U]) of the whole belief set as int32 tensors on `device`.
    """
    world, rank = group_size(group), group_rank(group)
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
    if guess_state is None:
        guess_state = [capacity // 2]               # no history: blocks at full capacity
    if positions is not None:
        pos = torch.as_tensor(positions).to(device=device, dtype=torch.int32)
        g_first, g_last = pos[first.long()], pos[last.long()]
    else:
        offset = rank * capacity
        g_first, g_last = first + offset, last + offset
    # Most beliefs share their tuple with others, so the blocks are first sized by a guess (twice the largest count seen in the
    # previous exchange); the header carries the true count, and if any rank overflowed -- every rank sees that in the gathered
    # headers, so the decision is consistent -- the exchange is repeated once at full capacity.
    guess = min(capacity, max(256, 2 * guess_state[0]))
    while True:
        k = min(u, guess)
        buf = torch.empty((guess + 1, w + 2), dtype=torch.int32, device=device)
        buf[0].fill_(u)
        buf[1:k + 1, :w] = tuples[:k]
        buf[1:k + 1, w] = g_first[:k]
        buf[1:k + 1, w + 1] = g_last[:k]
        gathered = torch.empty((world * (guess + 1), w + 2), dtype=torch.int32, device=device)
        all_gather_rows(gathered, buf, group)
        first_rows, last_rows, max_count = merge_fn(gathered, world, guess + 1, w)
        guess_state[0] = max_count
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
        self._guess = [256]             # largest record count of this instance's previous exchange (sizes the next blocks)
        self.world = group_size(group)
        self.rank = group_rank(group)
        self.last_payload_bytes = 0
        self.trace = False              # set True to record per-phase wall times (device-synchronised) in `last_phases`
        self.last_phases = {}

    def bounds(self, n_rows: int) -> tuple:
        self._fixed_cap = -(-n_rows // self.world)
        return shard_bounds(n_rows, self.world, self.rank)

    def backup(self, local_belief_set, value_function, append: bool = False, belief_dominance_prune: bool = False, positions=None,
               capacity: int | None = None):
        """`positions` / `capacity`: explicit positions of the local beliefs in the whole set and an upper bound on any rank's
        shard size (the sharded solve); default: contiguous blocks of `bounds()` / `set_capacity()`."""
        from .value_function import ValueFunction
        t = _PhaseTimer(self.model.device.device) if self.trace else None
        if self.exchange == 'tuples':
            dev = self.model.device
            tuples, first, last = self.solver.select_tuples_device(self.model, local_belief_set, value_function, belief_dominance_prune)
            if t: t.mark('local select')
            n_local = len(local_belief_set)
            assert positions is None or not belief_dominance_prune, 'explicit positions index the unfiltered local belief set'
            cap = int(capacity) if capacity is not None else self._capacity(n_local)
            g_tuples, _, g_last = exchange_tuples(tuples, first, last, cap, dev.device, self.group, merge_fn=dev.group_record_blocks,
                                                  positions=positions, guess_state=self._guess)
            self.last_payload_bytes = (min(cap, max(256, 2 * self._guess[0])) + 1) * (tuples.shape[1] + 2) * 4 * self.world
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
        all_reduce_(c, dist.ReduceOp.MAX, self.group)
        self._cap = (n_local, int(c[0]))
        return self._cap[1]

    def compute_change(self, value_function, new_value_function, local_belief_set) -> float:
        """max over all shards of the local change: one scalar all-reduce(max)."""
        local = self.solver.compute_change(value_function, new_value_function, local_belief_set)
        # np.max / torch.max propagate NaN (a NaN belief row makes the change NaN in the reference); MAX over ranks would drop it
        t = torch.tensor([local if local == local else float('inf'), 1.0 if local != local else 0.0], dtype=torch.float64,
                         device=self.model.device.device)
        all_reduce_(t, dist.ReduceOp.MAX, self.group)
        return float('nan') if float(t[1]) > 0 else float(t[0])


class ShardedSolveState:
    """
    The distributed half of `PBVI_Solver.solve(..., group=...)` (reference loop src/pomdp.py:2306-2389) -- one process per GPU:

      expand          runs on group rank 0 only (the reference's host RNG draws happen once) and the new belief rows are broadcast;
                      every rank then forms the same union, so the whole belief set is replicated (50 000 beliefs of the olfactory
                      model are 8.8 GB of 180 GB) and `expand_*` flavours that read all of it (SSEA, GER, SSGA, RA) work unchanged.
      ownership       of the rows that the union appended, rank r owns a contiguous slice (`shard_bounds` over the fresh rows).
                      A rank's owned rows only ever grow by appending, so its local BeliefSet keeps the lineage chain that makes
                      `compute_change` incremental; `positions` remembers where every owned row sits in the whole set.
      full backup     every rank selects the tuples of its owned rows; the tuple exchange carries explicit positions, the merge
                      orders by position, every rank assembles the merged set: the value function is the single-process one, bit
                      for bit and in order (first position, last action).
      new-points backup (FSVI / HSVI / Perseus: <= max_belief_growth rows): sharded the same way when there are at least
                      `replicate_below` rows per rank, otherwise every rank backs the few rows up itself -- same kernels on the same
                      inputs give the same bytes, and no collective is cheaper than one.
      compute_change  local maxima over the owned rows (incremental), one scalar all-reduce(max).
    """

    def __init__(self, solver, model, group=None, replicate_below: int = 64):
        self.solver, self.model, self.group = solver, model, group
        self.sb = ShardedBackup(solver, model, group)
        self.sb_new = ShardedBackup(solver, model, group)      # its own block-size history: the two backups differ in size
        self.world, self.rank = self.sb.world, self.sb.rank
        self.replicate_below = int(replicate_below)
        self.local_set = None
        self._rows = None            # [capacity, S] owned rows
        self._pos = None             # [capacity] int32 position of every owned row in the whole belief set
        self.n_local = 0
        self.n_global = 0
        self.max_owned = 0           # upper bound on any rank's number of owned rows
        self.stats = {'broadcast_rows': 0, 'sharded_backups': 0, 'replicated_backups': 0, 'sharded_pairs': 0.0}

    # ---- expansion on rank 0, rows broadcast ---------------------------------------------------------------------------
    def expand(self, model, belief_set, value_function, max_generation, **params):
        import time
        from .belief import BeliefSet
        dev = model.device.device
        t0 = time.perf_counter()
        new = None
        if self.rank == 0:
            new = self.solver.expand(model=model, belief_set=belief_set, value_function=value_function, max_generation=max_generation,
                                     **params)
            head = torch.tensor([len(new)], dtype=torch.int64, device=dev)
        else:
            head = torch.zeros((1,), dtype=torch.int64, device=dev)
        t1 = time.perf_counter()
        broadcast_(head, 0, self.group)
        n = int(head[0])
        if self.rank == 0:
            rows = new.belief_array.contiguous()
        else:
            rows = torch.empty((n, model.state_count), dtype=torch.float64, device=dev)
        if n:
            broadcast_(rows, 0, self.group)
        self.stats['broadcast_rows'] += n
        out = new if self.rank == 0 else BeliefSet(model, rows)
        t2 = time.perf_counter()
        self.stats['expand_rank0_s'] = self.stats.get('expand_rank0_s', 0.0) + (t1 - t0)
        self.stats['expand_broadcast_s'] = self.stats.get('expand_broadcast_s', 0.0) + (t2 - t1)     # includes waiting for rank 0
        return out

    # ---- ownership of the rows a union appended ---------------------------------------------------------------------------
    def absorb(self, belief_set) -> None:
        """Called after every `belief_set = belief_set.union(new)` (and once for the initial set): rows [n_global, len) are new."""
        import time
        from .belief import BeliefSet
        t0 = time.perf_counter()
        n_prev, n_now = self.n_global, len(belief_set)
        n_fresh = n_now - n_prev
        assert n_fresh >= 0, 'the belief set of a solve only grows'
        lo, hi = shard_bounds(n_fresh, self.world, self.rank)
        self.max_owned += -(-n_fresh // self.world)
        self.n_global = n_now
        k = hi - lo
        if k == 0 and self.local_set is not None:
            return
        S, dev = self.model.state_count, self.model.device.device
        n0, n1 = self.n_local, self.n_local + k
        if self._rows is None or self._rows.shape[0] < n1:
            cap = max(256, 2 * n1)
            rows = torch.empty((cap, S), dtype=torch.float64, device=dev)
            pos = torch.empty((cap,), dtype=torch.int32, device=dev)
            if n0:
                rows[:n0], pos[:n0] = self._rows[:n0], self._pos[:n0]
            self._rows, self._pos = rows, pos           # earlier local sets keep viewing the old store
        if k:
            self._rows[n0:n1] = belief_set.belief_array[n_prev + lo:n_prev + hi]
            self._pos[n0:n1] = torch.arange(n_prev + lo, n_prev + hi, dtype=torch.int32, device=dev)
        grown = BeliefSet(self.model, self._rows[:n1])
        if self.local_set is not None:
            grown._inherit(self.local_set, n0)
        self.local_set, self.n_local = grown, n1
        self.stats['absorb_s'] = self.stats.get('absorb_s', 0.0) + (time.perf_counter() - t0)

    # ---- backups ------------------------------------------------------------------------------------------------------
    def backup_full(self, value_function):
        """Backup over the whole belief set (full_backup=True flavours), sharded by ownership."""
        self.stats['sharded_backups'] += 1
        self.stats['sharded_pairs'] += float(self.n_global) * len(value_function)
        return self.sb.backup(self.local_set, value_function, append=False, positions=self._pos[:self.n_local],
                              capacity=max(self.max_owned, 1))

    def backup_new(self, new_belief_set, value_function):
        """New-points backup (append=True) over the expansion's rows: contiguous shards, or replicated when they are few."""
        from .belief import BeliefSet
        n = len(new_belief_set)
        if n < self.replicate_below * self.world:
            self.stats['replicated_backups'] += 1
            return self.solver.backup(self.model, new_belief_set, value_function, append=True, belief_dominance_prune=False)
        self.stats['sharded_backups'] += 1
        self.stats['sharded_pairs'] += float(n) * len(value_function)
        lo, hi = shard_bounds(n, self.world, self.rank)
        self.sb_new.set_capacity(-(-n // self.world))
        return self.sb_new.backup(BeliefSet(self.model, new_belief_set.belief_array[lo:hi]), value_function, append=True)

    def compute_change(self, value_function, new_value_function) -> float:
        return self.sb.compute_change(value_function, new_value_function, self.local_set)

    # ---- the value-function size limiter draws on the host RNG: rank 0 decides, everyone applies ------------------------------
    def limit_value_function(self, value_function, belief_set, max_belief_growth):
        from .value_function import ValueFunction
        dev = self.model.device.device
        head = torch.zeros((2,), dtype=torch.int64, device=dev)
        keep = None
        if self.rank == 0:
            limited, n_useful = self.solver._limit_value_function(self.model, value_function, belief_set, max_belief_growth, return_keep=True)
            keep = torch.as_tensor(limited, dtype=torch.int64, device=dev)
            head[0], head[1] = keep.shape[0], n_useful
        broadcast_(head, 0, self.group)
        if self.rank != 0:
            keep = torch.empty((int(head[0]),), dtype=torch.int64, device=dev)
        broadcast_(keep, 0, self.group)
        k = keep.cpu().numpy()
        rows = value_function.alpha_vector_array[keep]
        return ValueFunction(self.model, rows, value_function.actions[k], _trusted=True, _hashes=value_function.row_hashes[k]), int(head[1])
