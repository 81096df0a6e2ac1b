"""
CPU-only checks of the drop-in boundary and the host logic:
  * libpbvi_b200.so loads without a GPU and exports exactly the entry points include/pbvi_b200.h declares,
    and the ctypes table of the binding covers every one of them;
  * the error convention (negative status + pbvi_last_error) works without touching a device;
  * host-side set bookkeeping (dict-insertion order on 128-bit row keys, first position / last action) and the belief sharding
    bounds behave as the reference's dict semantics require;
  * the host half of the sparse-row transport (pbvi_pack_rows_host / pbvi_pack_slabs_host: plain C++, no device) and the host
    twin of the sharded merge reproduce bytes / dict order;
  * the product package never imports the oracle.
No DEVICE entry point is called here.
"""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, 'include', 'pbvi_b200.h')


def declared_symbols():
    text = open(HEADER).read()
    return sorted(set(re.findall(r'PBVI_API\s+[\w\s\*]+?\b(pbvi_\w+)\s*\(', text)))


@pytest.fixture(scope='module')
def lib_path():
    from pomdp_pbvi_exploration_b200.build import build_library
    return build_library()


def test_header_declares_the_expected_surface():
    syms = declared_symbols()
    for must in ['pbvi_model_create', 'pbvi_model_destroy', 'pbvi_backup', 'pbvi_backup_host', 'pbvi_backup_select', 'pbvi_backup_assemble',
                 'pbvi_max_values', 'pbvi_belief_update', 'pbvi_belief_successors', 'pbvi_row_hash', 'pbvi_rows_equal', 'pbvi_vi_sweep',
                 'pbvi_last_error']:
        assert must in syms


def test_library_exports_every_declared_symbol(lib_path):
    out = subprocess.run(['nm', '-D', '--defined-only', lib_path], capture_output=True, text=True, check=True).stdout
    exported = sorted(set(re.findall(r'\bT (pbvi_\w+)', out)))
    assert exported == declared_symbols()


def test_binding_table_matches_header(lib_path):
    from pomdp_pbvi_exploration_b200 import _native
    lib = _native.load_library()
    table = set(_native.SIGNATURES) | {'pbvi_last_error'}
    assert table == set(declared_symbols())
    for name in table:
        assert hasattr(lib, name)
    assert lib.pbvi_version() >= 100


def test_error_convention_without_a_device(lib_path):
    from pomdp_pbvi_exploration_b200 import _native
    lib = _native.load_library()
    assert lib.pbvi_model_dims(None, None, None, None, None) == _native.PBVI_ERR_BAD_ARG
    assert b'NULL' in lib.pbvi_last_error()
    h = ctypes.c_void_p()
    rc = lib.pbvi_model_create(0, 1, 1, 1, None, None, None, None, 0, ctypes.byref(h))
    assert rc == _native.PBVI_ERR_BAD_ARG and h.value is None
    with pytest.raises(ValueError):
        _native._check(rc)


def test_comm_and_small_path_entry_points_reject_bad_arguments(lib_path):
    """The multi-GPU collectives (pbvi_comm_*) and the small-model backup are part of the C ABI: exported, int status, bad-argument
    errors without touching a device or NCCL."""
    from pomdp_pbvi_exploration_b200 import _native
    lib = _native.load_library()
    h = ctypes.c_void_p()
    buf = (ctypes.c_ubyte * 128)()
    assert lib.pbvi_comm_init(None, buf, 0, 2, ctypes.byref(h)) == _native.PBVI_ERR_BAD_ARG and h.value is None
    assert lib.pbvi_comm_rank(None, None, None) == _native.PBVI_ERR_BAD_ARG
    assert lib.pbvi_comm_destroy(None) == _native.PBVI_OK
    assert lib.pbvi_allgather_tuples(None, None, 4, 5, None, None) == _native.PBVI_ERR_BAD_ARG
    assert lib.pbvi_allreduce_max(None, None, 1, None) == _native.PBVI_ERR_BAD_ARG
    assert lib.pbvi_broadcast_rows(None, None, 0, 0, None) == _native.PBVI_ERR_BAD_ARG
    assert lib.pbvi_backup_small_eligible(None, 10, 10) == 0
    n = ctypes.c_int(7)
    assert lib.pbvi_backup_small(None, None, 1, None, 1, 0.9, None, None, None, ctypes.byref(n), None) == _native.PBVI_ERR_BAD_ARG
    assert lib.pbvi_perseus_walk(None, None, None, None, 3, None, None, None) == _native.PBVI_ERR_BAD_ARG


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, 'pomdp_pbvi_exploration_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh')):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle\b', text, flags=re.M), f


def test_group_by_key_is_dict_insertion_order():
    from pomdp_pbvi_exploration_b200.sets import group_by_key
    rng = np.random.default_rng(0)
    keys = rng.integers(-5, 5, (200, 2)).astype(np.int64)
    first, last, inverse = group_by_key(keys)
    table = {}
    for i, k in enumerate(map(tuple, keys.tolist())):
        if k in table:
            table[k][1] = i
        else:
            table[k] = [i, i]
    assert first.tolist() == [v[0] for v in table.values()]
    assert last.tolist() == [v[1] for v in table.values()]
    assert all(tuple(keys[first[g]]) == tuple(keys[i]) for i, g in enumerate(inverse))
    e = group_by_key(np.zeros((0, 2), dtype=np.int64))
    assert all(x.shape == (0,) for x in e)


def test_unique_rows_first_packed_and_wide():
    import importlib
    solver = importlib.import_module('pomdp_pbvi_exploration_b200.solver')
    rng = np.random.default_rng(1)
    for width, hi in [(4, 1000), (23, 10 ** 6)]:           # packs into one int64 / falls back to row-wise unique
        keys = rng.integers(0, hi, (50, width))
        keys = np.concatenate([keys, keys[::3], keys[:5]])
        first, last, inverse = solver.unique_rows_first(keys)
        table = {}
        for i, k in enumerate(map(tuple, keys.tolist())):
            table.setdefault(k, [i, i])[1] = i
        assert first.tolist() == [v[0] for v in table.values()] and last.tolist() == [v[1] for v in table.values()]
        assert np.array_equal(keys[first][inverse], keys)


@pytest.mark.parametrize('unit', [1, 64, 448, 960, 1024])
def test_streamed_chunk_plan_covers_the_rows_in_order(unit):
    """Chunk plan of the streamed (host-buffer) select: contiguous, in order, every boundary but the last a multiple of the ship unit,
    sizes double from the first chunk up to the cap, and a tail shorter than the first chunk joins its predecessor."""
    from pomdp_pbvi_exploration_b200.solver import PBVI_Solver
    s = PBVI_Solver(expand_function='perseus')
    for n in (1, 63, 2048, 2100, 2500, 4000, 6250, 10000, 50000):
        plan = s._chunk_plan(n, unit)
        assert plan[0][0] == 0 and plan[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(plan, plan[1:]))
        assert all(lo < hi for lo, hi in plan)
        assert all(hi % unit == 0 for _, hi in plan[:-1])
        first = max(1, s.STREAM_FIRST_CHUNK // unit) * unit
        cap = max(first, int(round(s.STREAM_CHUNK / unit)) * unit)
        sizes = [hi - lo for lo, hi in plan]
        assert all(x <= cap for x in sizes[:-1]) and sizes[-1] < cap + first
        assert all(b in (min(cap, 2 * a), n - sum(sizes[:i + 1])) or i + 2 == len(sizes) for i, (a, b) in enumerate(zip(sizes, sizes[1:])))
        if len(plan) > 1:
            assert sizes[0] == first and sizes[-1] >= first


def test_shard_bounds_cover_rows_contiguously():
    from pomdp_pbvi_exploration_b200.parallel import shard_bounds
    for n in [0, 1, 7, 8, 9, 50000]:
        for world in [1, 2, 3, 8]:
            spans = [shard_bounds(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            assert all(hi - lo <= -(-n // world) for lo, hi in spans)


def test_model_constructor_matches_reference_tensors():
    """The package's Model derives the same tensors as the reference's constructor (golden fixtures of the small models)."""
    from conftest import load_golden
    from pomdp_pbvi_exploration_b200 import Model
    for tag in ['tiger', 'grid4x4', 'hallway']:
        m = load_golden('model_' + tag)
        S, A, O = m['rto'].shape[0], m['rto'].shape[1], m['rto'].shape[2]
        model = Model(states=S, actions=A, observations=O, transitions=m['transition_table'], rewards=m['reward_table'],
                      observation_table=m['obs_table'], start_probabilities=m['start'])
        assert np.array_equal(model.reachable_states, m['reach'])
        assert np.array_equal(model.reachable_probabilities, m['probs'])
        assert np.array_equal(model.reachable_transitional_observation_table, m['rto'])
        assert np.array_equal(model.expected_rewards_table, m['rbar'])


def test_olfactory_recipe_matches_reference_model():
    from conftest import load_golden
    from pomdp_pbvi_exploration_b200.recipes import olfactory_wrap_model
    m = load_golden('model_olfactory_wrap')
    model = olfactory_wrap_model()
    assert np.array_equal(model.reachable_states, m['reach'])
    assert np.array_equal(model.reachable_transitional_observation_table, m['rto'])
    assert np.array_equal(model.expected_rewards_table, m['rbar'])
    assert np.array_equal(model.start_probabilities, m['start'])


def _unpack_numpy(bitmap, row_start, packed, n, L):
    """Reference unpacking of the sparse-row transport format (what pbvi_unpack_rows does on the device)."""
    n_c = -(-L // 4)
    out = np.zeros((n, n_c * 4))
    chunks = packed.reshape(-1, 4)
    for i in range(n):
        bits = np.unpackbits(bitmap[i].view(np.uint8), bitorder='little')[:n_c].astype(bool)
        assert bits.sum() == row_start[i + 1] - row_start[i]
        out[i].reshape(n_c, 4)[bits] = chunks[row_start[i]:row_start[i + 1]]
    return out[:, :L]


@pytest.mark.parametrize('n,L', [(0, 5), (1, 1), (7, 6), (100, 257), (33, 22021)])
def test_host_row_packer_round_trip(lib_path, n, L):
    """pbvi_pack_rows_host / pbvi_pack_slabs_host are pure host code (no device needed): bitmap + non-zero 4-double chunks, and
    the packed form unpacks to the same BYTES (-0.0, NaN payloads, ragged last chunk, all-zero and dense rows)."""
    import ctypes
    lib = ctypes.CDLL(lib_path)
    rng = np.random.default_rng(n + L)
    rows = np.zeros((n, L))
    for i in range(n):
        if i % 4 == 0:
            lo = int(rng.integers(0, L)); k = int(rng.integers(0, L - lo + 1))
            rows[i, lo:lo + k] = rng.random(k)
        elif i % 4 == 1:
            rows[i, rng.choice(L, min(L, 5), replace=False)] = rng.random(min(L, 5))
        elif i % 4 == 2:
            rows[i] = rng.random(L)
    if n > 1:
        rows[0, -1] = -0.0
        rows[1, 0] = np.nan
    n_c = -(-L // 4); W = -(-n_c // 32)
    bm = np.zeros((max(n, 1), W), dtype=np.uint32)
    rs = np.zeros(n + 1, dtype=np.int32)
    pk = np.zeros(n * n_c * 4 + 4)
    total = ctypes.c_int64(-1)
    rc = lib.pbvi_pack_rows_host(ctypes.c_void_p(rows.ctypes.data), n, L, ctypes.c_void_p(bm.ctypes.data), ctypes.c_void_p(rs.ctypes.data),
                                 ctypes.c_void_p(pk.ctypes.data), ctypes.byref(total))
    assert rc == 0 and total.value == rs[n]
    got = _unpack_numpy(bm, rs, pk[:total.value * 4], n, L)
    assert np.array_equal(got.view(np.uint64), rows.view(np.uint64))
    if n == 0:
        return
    # slabs of 3 rows, two "threads" (slabs 0,2,4,.. and 1,3,5,..): same bytes, counts published per slab
    SL = 3
    n_slabs = -(-n // SL)
    region = SL * n_c * 4 + 4
    bm2 = np.zeros((n, W), dtype=np.uint32)
    rs2 = np.zeros((n_slabs, SL + 1), dtype=np.int32)
    pk2 = np.zeros(n_slabs * region)
    totals = np.full(n_slabs, -1, dtype=np.int64)
    lib.pbvi_pack_slabs_host.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                         ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p]
    for t in range(2):
        assert lib.pbvi_pack_slabs_host(rows.ctypes.data, n, L, SL, t, 2, bm2.ctypes.data, rs2.ctypes.data, pk2.ctypes.data, region,
                                        totals.ctypes.data) == 0
    assert np.all(totals >= 0)
    for i in range(n_slabs):
        r0, r1 = i * SL, min(n, (i + 1) * SL)
        got = _unpack_numpy(bm2[r0:r1], rs2[i], pk2[i * region:i * region + totals[i] * 4], r1 - r0, L)
        assert np.array_equal(got.view(np.uint64), rows[r0:r1].view(np.uint64))


def test_merge_blocks_host_matches_a_dict_over_all_records():
    """Host twin of pbvi_group_record_blocks (used by the gloo tests of the tuple exchange): grouping of the valid records of the
    gathered blocks == one dict over the records visited in order of their first-position word (the position of the belief in the
    whole set, which need not follow the buffer order); `last` = record with the largest last-position."""
    import torch
    from pomdp_pbvi_exploration_b200.parallel import merge_blocks_host
    rng = np.random.default_rng(2)
    world, block_rows, w = 3, 9, 2
    blocks = rng.integers(0, 99, (world, block_rows, w + 2)).astype(np.int32)
    counts = [8, 0, 5]
    pool = rng.integers(0, 4, (6, w))
    firsts = rng.permutation(1000)                                # unique first positions, interleaved across the ranks
    for r in range(world):
        blocks[r, 0, :] = counts[r]
        blocks[r, 1:1 + counts[r], :w] = pool[rng.integers(0, 6, counts[r])]
        blocks[r, 1:1 + counts[r], w] = firsts[r * 10:r * 10 + counts[r]]
        blocks[r, 1:1 + counts[r], w + 1] = rng.permutation(100)[:counts[r]] + 100 * r
    flat = blocks.reshape(-1, w + 2)
    first, last, mx = merge_blocks_host(torch.as_tensor(flat), world, block_rows, w)
    rows = [r * block_rows + j for r in range(world) for j in range(1, 1 + counts[r])]
    table = {}
    for row in sorted(rows, key=lambda x: flat[x, w]):
        key = flat[row, :w].tobytes()
        if key not in table:
            table[key] = [row, row]
        elif flat[row, w + 1] > flat[table[key][1], w + 1]:
            table[key][1] = row
    assert mx == 8
    assert first.tolist() == [v[0] for v in table.values()] and last.tolist() == [v[1] for v in table.values()]
