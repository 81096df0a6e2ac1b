"""
World-size-2/3 test of the N > 1 path on CPU (gloo): the key-first exchange of new alpha rows (`exchange_new_rows`) must give every rank the value function a single process computes over the whole belief set
(first position, last action).  The per-rank "backup" here is a stand-in that only exercises the exchange: rows are
produced by the oracle, keys by a host hash, and byte-equality by torch on CPU.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import pbvi_oracle as orc


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _host_hash(rows: np.ndarray) -> np.ndarray:
    import hashlib
    out = np.zeros((rows.shape[0], 2), dtype=np.int64)
    for i, r in enumerate(rows):
        out[i] = np.frombuffer(hashlib.blake2b(r.tobytes(), digest_size=16).digest(), dtype=np.int64)
    return out


def _rows_equal(ra, ia, rb, ib):
    return torch.tensor([int(torch.equal(ra[int(i)], rb[int(j)])) for i, j in zip(ia, ib)], dtype=torch.int32)


def _worker(rank, world, port, rows_all, actions_all, counts, result_queue):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from pomdp_pbvi_exploration_b200.parallel import exchange_new_rows, shard_bounds
    lo, hi = shard_bounds(rows_all.shape[0], world, rank)
    # local dedup (what the per-rank backup returns): first position, last action
    local_rows, local_actions, _ = orc.dedup_rows(rows_all[lo:hi], actions_all[lo:hi])
    local_hash = _host_hash(local_rows)
    merged_rows, merged_actions, merged_hashes, payload = exchange_new_rows(torch.as_tensor(local_rows), local_actions, local_hash, _rows_equal)
    assert np.array_equal(merged_hashes, _host_hash(merged_rows.numpy()))
    result_queue.put((rank, merged_rows.numpy(), merged_actions, payload))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize('n_rows,world', [(37, 2), (5, 2), (1, 2), (41, 3)])
def test_sharded_merge_equals_single_process(n_rows, world):
    rng = np.random.default_rng(n_rows)
    S = 11
    base = rng.random((6, S))
    pick = rng.integers(0, 6, n_rows)
    rows_all = base[pick]                               # many duplicates, spread over both shards
    actions_all = rng.integers(0, 4, n_rows).astype(np.int64)
    want_rows, want_actions, _ = orc.dedup_rows(rows_all, actions_all)
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, rows_all, actions_all, None, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, rows, actions, payload in results:
        assert np.array_equal(rows, want_rows), rank
        assert np.array_equal(actions, want_actions), rank
        assert payload > 0


def _tuple_worker(rank, world, port, keys_all, result_queue):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from pomdp_pbvi_exploration_b200.parallel import exchange_tuples, shard_bounds
    from pomdp_pbvi_exploration_b200.sets import unique_rows_first
    lo, hi = shard_bounds(keys_all.shape[0], world, rank)
    local = keys_all[lo:hi]
    first, last, _ = unique_rows_first(local)
    cap = -(-keys_all.shape[0] // world)
    # block-size history of 100 records -> first guess 256: the 1300-belief case overflows it and must retry consistently
    g_tuples, g_first, g_last = exchange_tuples(local[first], first, last, cap, torch.device('cpu'), guess_state=[100])
    result_queue.put((rank, g_tuples.numpy(), g_first.numpy(), g_last.numpy()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize('n_beliefs,world', [(50, 2), (3, 2), (64, 3), (1300, 2)])
def test_tuple_exchange_equals_single_process(n_beliefs, world):
    """The compact exchange: distinct (a*, v*) tuples + first/last belief positions of every shard, merged identically on every
    rank, must equal the distinct tuples (first occurrence order, first / last positions) of the whole belief set."""
    from pomdp_pbvi_exploration_b200.sets import unique_rows_first
    rng = np.random.default_rng(n_beliefs)
    if n_beliefs > 1000:
        # more distinct tuples per rank than the first-guess block size (256 records): the exchange must notice the overflow in the
        # gathered headers on every rank and repeat once at full capacity
        pool = np.unique(rng.integers(0, 40, (900, 4)), axis=0)
        keys_all = pool[rng.integers(0, pool.shape[0], n_beliefs)]
    else:
        pool = rng.integers(0, 40, (7, 4))
        keys_all = pool[rng.integers(0, 7, n_beliefs)]
    f, l, _ = unique_rows_first(keys_all)
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_tuple_worker, args=(r, world, port, keys_all, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, g_tuples, g_first, g_last in results:
        assert np.array_equal(g_tuples, keys_all[f]) and np.array_equal(g_first, f) and np.array_equal(g_last, l), rank


def _owned_positions(n_total, world, rank, growth):
    """Positions owned by `rank` under the sharded solve's append-only ownership: of every batch of `growth` appended rows a rank
    owns one contiguous slice (ShardedSolveState.absorb)."""
    from pomdp_pbvi_exploration_b200.parallel import shard_bounds
    pos, start = [], 0
    while start < n_total:
        n = min(growth, n_total - start)
        lo, hi = shard_bounds(n, world, rank)
        pos.extend(range(start + lo, start + hi))
        start += n
    return np.array(pos, dtype=np.int64)


def _interleaved_worker(rank, world, port, keys_all, growth, result_queue):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from pomdp_pbvi_exploration_b200.parallel import exchange_tuples
    from pomdp_pbvi_exploration_b200.sets import unique_rows_first
    pos = _owned_positions(keys_all.shape[0], world, rank, growth)
    local = keys_all[pos]
    first, last, _ = unique_rows_first(local) if pos.shape[0] else (np.zeros(0, dtype=np.int64),) * 3
    cap = max(1, max(_owned_positions(keys_all.shape[0], world, r, growth).shape[0] for r in range(world)))
    g_tuples, g_first, g_last = exchange_tuples(local[first].reshape(-1, keys_all.shape[1]), first, last, cap, torch.device('cpu'),
                                                positions=pos.astype(np.int32), guess_state=[100])
    result_queue.put((rank, g_tuples.numpy(), g_first.numpy(), g_last.numpy()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize('n_beliefs,world,growth', [(50, 2, 7), (64, 3, 10), (5, 3, 1), (1300, 2, 100)])
def test_tuple_exchange_with_interleaved_ownership(n_beliefs, world, growth):
    """The sharded solve's ownership interleaves the ranks' rows in the whole belief set; the exchange then carries explicit
    positions and the merge must still return the single-process grouping: tuples in order of first position, first / last positions."""
    from pomdp_pbvi_exploration_b200.sets import unique_rows_first
    rng = np.random.default_rng(n_beliefs + growth)
    pool = np.unique(rng.integers(0, 40, (900 if n_beliefs > 1000 else 9, 4)), axis=0)
    keys_all = pool[rng.integers(0, pool.shape[0], n_beliefs)]
    f, l, _ = unique_rows_first(keys_all)
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_interleaved_worker, args=(r, world, port, keys_all, growth, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, g_tuples, g_first, g_last in results:
        assert np.array_equal(g_tuples, keys_all[f]) and np.array_equal(g_first, f) and np.array_equal(g_last, l), rank
