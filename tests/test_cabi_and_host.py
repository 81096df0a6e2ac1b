"""
CPU-only checks of the drop-in boundary and the host logic:
  * libpbvi_b200.so loads without a GPU and exports exactly the entry points include/pbvi_b200.h declares,
    and the ctypes table of the binding covers every one of them;
  * the error convention (negative status + pbvi_last_error) works without touching a device;
  * host-side set bookkeeping (dict-insertion order on 128-bit row keys, first position / last action) and the belief sharding
    bounds behave as the reference's dict semantics require;
  * the product package never imports the oracle.
No compute entry point is called here.
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
