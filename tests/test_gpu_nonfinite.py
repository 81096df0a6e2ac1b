"""
GPU parity tests for NON-FINITE inputs (np.argmax / NaN semantics of the reference, src/pomdp.py:1494-1506).

`Belief.update(a, o)` returns an all-NaN row for an impossible observation (0/0, src/pomdp.py:405-411) and nothing in the
reference stops such a row from entering a belief set; a value function can hold +-inf (e.g. an `-inf` initial lower bound).
NumPy's arithmetic then yields NaN scores, and `np.argmax` treats a NaN as the maximum and returns the FIRST one.  The engine
must (1) never index out of bounds, (2) return what the reference returns: a row that holds a non-finite belief entry gets
v* = 0 for every (a, o) and a* = 0; with non-finite alphas every zero-skipping rule is switched off and the dense product is
computed.  (compute-sanitizer is closed on the GPU pool: the kernels clamp every gathered index instead, and `test_assemble_clamps_bad_indices`
feeds them indices that would fault without the clamp.)
"""
import numpy as np
import pytest

from conftest import load_golden
from oracle import pbvi_oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def torch_cuda():
    import torch
    assert torch.cuda.is_available(), 'these tests need a CUDA device'
    return torch


_cache = {}


def _model(tag):
    from pomdp_pbvi_exploration_b200._native import DeviceModel
    if tag not in _cache:
        m = load_golden('model_' + tag)
        reach = m['reach'].astype(np.int64)
        probs = m['probs'] if 'probs' in m else np.full(reach.shape, 1.0 / reach.shape[2])
        _cache[tag] = (DeviceModel(reach, probs, m['rto'], m['rbar']), m, reach)
    return _cache[tag]


def _nan_successor(dev, reach, rto, beliefs):
    """A NaN row made the way the reference makes one: Belief.update on an observation of probability zero."""
    A, O = rto.shape[1], rto.shape[2]
    for i in range(beliefs.shape[0]):
        for a in range(A):
            for o in range(O):
                want = orc.belief_update(reach, rto, beliefs[i], a, o)
                if np.isnan(want).all():
                    out, mass = dev.belief_update(beliefs[i:i + 1], [a], [o])
                    got = out.cpu().numpy()[0]
                    assert np.isnan(got).all() and float(mass[0]) == 0.0
                    return got
    pytest.skip('model has no impossible observation')


@pytest.mark.parametrize('tag', ['grid4x4_noloop', 'olfactory_wrap', 'tigergrid'])
def test_backup_with_nan_belief_row(torch_cuda, tag):
    dev, m, reach = _model(tag)
    g = load_golden('backup_' + tag)
    gamma = float(m['gamma'])
    B = g['beliefs'][:24].copy()
    V = g['alphas'][:40]
    nan_row = _nan_successor(dev, reach, m['rto'], B)
    half = B[3].copy()
    half[1] = np.nan                                     # a single NaN entry poisons the whole row in the reference as well
    B = np.concatenate([B[:5], nan_row[None], B[5:17], half[None], B[17:]])
    alpha, act, vstar, value = [t.cpu().numpy() for t in dev.backup(B, V, gamma)]
    with np.errstate(all='ignore'):
        ref = orc.backup(reach, m['rto'], m['rbar'], gamma, B, V)
    bad = np.isnan(B).any(axis=1)
    assert bad.sum() == 2
    # the reference: every score of such a row is NaN -> v* = 0 everywhere, every value NaN -> a* = 0, the row of tuple (0, [0..0])
    assert np.all(ref['v_star'][bad] == 0) and np.all(ref['a_star'][bad] == 0)
    assert np.all(vstar[bad] == 0) and np.all(act[bad] == 0) and np.isnan(value[bad]).all()
    exact = reach.shape[2] == 1
    if exact:
        assert np.array_equal(alpha[bad], ref['alpha'][bad])
    else:
        np.testing.assert_allclose(alpha[bad], ref['alpha'][bad], rtol=1e-9, atol=1e-12)
    # the finite rows of the same call are untouched by their neighbour (same 64-belief tile, same 16-belief row group)
    clean = dev.backup(B[~bad], V, gamma)
    assert np.array_equal(alpha[~bad], clean[0].cpu().numpy()) and np.array_equal(act[~bad], clean[1].cpu().numpy())
    assert np.array_equal(vstar[~bad], clean[2].cpu().numpy())
    # max_v b.alpha_v (compute_change, simulations): NaN maximum, index 0 (np.max / np.argmax)
    mx, arg = dev.max_values(B, V)
    mx, arg = mx.cpu().numpy(), arg.cpu().numpy()
    assert np.isnan(mx[bad]).all() and np.all(arg[bad] == 0)
    want_mx, want_arg = orc.max_values(B[~bad], V)
    np.testing.assert_allclose(mx[~bad], want_mx, rtol=1e-12, atol=1e-15)


@pytest.mark.parametrize('tag', ['grid4x4_noloop', 'tiger', 'synth300'])
@pytest.mark.parametrize('poison', ['-inf', 'nan', '+inf', 'mixed'])
def test_backup_with_nonfinite_alphas(torch_cuda, tag, poison):
    dev, m, reach = _model(tag)
    g = load_golden('backup_' + tag)
    gamma = float(m['gamma'])
    B = g['beliefs'][:40]
    rng = np.random.default_rng(12)
    S = g['alphas'].shape[1]
    lo, hi = float(np.min(g['alphas'])), float(np.max(g['alphas']))
    V = np.concatenate([g['alphas'], lo + (hi - lo + 1.0) * rng.random((12, S))])[:12].copy()
    if poison == '-inf':
        V[3, S // 3] = -np.inf
    elif poison == 'nan':
        V[5, S // 2] = np.nan
    elif poison == '+inf':
        V[2, 0] = np.inf
    else:
        V[1, S - 1] = -np.inf
        V[7, S // 2] = np.inf
        V[9, :] = -np.inf                                  # an "unreachable" lower-bound vector
    alpha, act, vstar, value = [t.cpu().numpy() for t in dev.backup(B, V, gamma)]
    with np.errstate(all='ignore'):
        ref = orc.backup(reach, m['rto'], m['rbar'], gamma, B, V, return_scores=True)
    sc = ref['scores']
    # columns whose score is NaN in the reference are NaN here, so the first-NaN rule gives the same v*; rows without a NaN follow
    # the usual contract (identical where the top-2 gap is decided)
    has_nan = np.isnan(sc).any(axis=3)
    assert np.array_equal(vstar[has_nan], ref['v_star'][has_nan])
    fin = ~has_nan
    if fin.any():
        with np.errstate(all='ignore'):
            top = np.sort(sc, axis=3)
            gap = top[..., -1] - top[..., -2]
            decided = fin & ~(gap <= 1e-9 * np.maximum(1.0, np.abs(top[..., -1])))      # inf - inf = NaN counts as undecided below
            decided &= np.isfinite(gap) | (gap == np.inf)
        assert np.array_equal(vstar[decided], ref['v_star'][decided])
    assert vstar.min() >= 0 and vstar.max() < V.shape[0]
    vn = np.isnan(ref['values'])
    assert np.array_equal(np.isnan(value), vn)
    row_nan = vn.any(axis=1)
    assert np.array_equal(act[row_nan], ref['a_star'][row_nan])
    same = (act == ref['a_star']) & np.all(np.take_along_axis(vstar, act[:, None, None].astype(np.int64), axis=1)[:, 0, :] ==
                                            np.take_along_axis(ref['v_star'], ref['a_star'][:, None, None], axis=1)[:, 0, :], axis=1)
    assert same.mean() > 0.5
    if reach.shape[2] == 1:
        assert np.array_equal(alpha[same], ref['alpha'][same], equal_nan=True)
    else:
        np.testing.assert_allclose(alpha[same], ref['alpha'][same], rtol=1e-9, atol=1e-12, equal_nan=True)


def test_solver_backup_survives_nan_rows(torch_cuda):
    """PBVI_Solver.backup (select -> tuple grouping -> assemble -> byte-dedup) on a belief set with NaN rows: the rows of the
    finite beliefs are those of the clean set, plus the one row the reference derives for a NaN belief: tuple (a* = 0, v* = 0)."""
    import torch
    from pomdp_pbvi_exploration_b200 import BeliefSet, PBVI_Solver, ValueFunction
    from pomdp_pbvi_exploration_b200.recipes import olfactory_wrap_model
    model = olfactory_wrap_model(points_per_unit=6)
    dev = model.device
    solver = PBVI_Solver(gamma=0.99, eps=1e-6, expand_function='perseus')
    np.random.seed(3)
    from pomdp_pbvi_exploration_b200 import Belief
    bs = solver.expand_perseus(model, Belief(model), max_generation=60)
    vf = ValueFunction(model, model.expected_rewards_table.T, model.actions)
    for _ in range(4):
        vf = solver.backup(model, bs, vf, append=True, belief_dominance_prune=False)
    rows = bs.belief_array.clone()
    succ, mass = dev.belief_successors(rows[:8])
    nan_rows = succ.reshape(-1, rows.shape[1])[(mass.reshape(-1) == 0)][:3]
    assert nan_rows.shape[0] == 3 and bool(torch.isnan(nan_rows).all())
    dirty = BeliefSet(model, torch.cat([rows[:10], nan_rows[:1], rows[10:], nan_rows[1:]]))
    out_dirty = solver.backup(model, dirty, vf, append=False, belief_dominance_prune=False)
    out_clean = solver.backup(model, BeliefSet(model, rows), vf, append=False, belief_dominance_prune=False)
    O = model.observation_count
    extra = dev.backup_assemble(vf.alpha_vector_array, 0.99, [0], np.zeros((1, O), dtype=np.int32))
    a, b = out_dirty.alpha_vector_array.cpu().numpy(), out_clean.alpha_vector_array.cpu().numpy()
    keys_d = {r.tobytes() for r in a}
    keys_c = {r.tobytes() for r in b} | {extra.cpu().numpy()[0].tobytes()}
    assert keys_d == keys_c
    # compute_change over a set with NaN rows is NaN in the reference (np.max propagates); no crash here either
    ch = solver.compute_change(vf, out_dirty, dirty)
    assert np.isnan(ch)


def test_assemble_clamps_bad_indices(torch_cuda):
    """Tuples with out-of-range alpha / action indices (a caller error) are clamped, never dereferenced out of bounds."""
    dev, m, reach = _model('grid4x4_noloop')
    g = load_golden('backup_grid4x4_noloop')
    V = g['alphas']
    nV, O, A = V.shape[0], m['rto'].shape[2], m['rto'].shape[1]
    n = 70                                                   # grouped kernel path (n >= 32)
    rng = np.random.default_rng(0)
    acts = rng.integers(0, A, n).astype(np.int32)
    vsel = rng.integers(0, nV, (n, O)).astype(np.int32)
    good = dev.backup_assemble(V, 0.95, acts, vsel).cpu().numpy()
    acts2, vsel2 = acts.copy(), vsel.copy()
    vsel2[5, 0] = 0x7fffffff
    vsel2[9, 1] = -7
    acts2[11] = A + 100
    bad = dev.backup_assemble(V, 0.95, acts2, vsel2).cpu().numpy()
    keep = np.ones(n, dtype=bool)
    keep[[5, 9, 11]] = False
    assert np.array_equal(bad[keep], good[keep]) and np.isfinite(bad).all()
    few = dev.backup_assemble(V, 0.95, acts2[:12], vsel2[:12]).cpu().numpy()        # generic kernel path (n < 32)
    assert np.array_equal(few, bad[:12])
