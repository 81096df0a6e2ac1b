"""
Pins oracle/pbvi_oracle.py against outputs of the unmodified reference (tests/golden/*.npz, produced by
tests/golden/make_golden.py).  CPU only.  Tolerances:
  * R == 1 models (olfactory, 4x4-no_loop): alpha rows, deduped value functions, belief updates bit-exact.
  * R > 1 models: NumPy's einsum r-reduction order is implementation defined, alpha rows
    within 1e-12 relative here (the contract is 1e-9); argmax indices exact wherever the gap exceeds 1e-9.
"""
import numpy as np
import pytest

from oracle import pbvi_oracle as orc
from conftest import load_golden

MODELS = ['tiger', 'grid4x4', 'grid4x4_noloop', 'tigergrid', 'hallway', 'cheese', 'grid4x3', 'cit', 'synth300', 'olfactory_wrap']
GAP_TOL = 1e-9


def _model(tag):
    return load_golden('model_' + tag)


@pytest.mark.parametrize('tag', [t for t in MODELS if t != 'olfactory_wrap'])
def test_model_tensors(tag):
    m = _model(tag)
    rto = orc.build_rto(m['reach'], m['probs'], m['obs_table'])
    assert np.array_equal(rto, m['rto'])
    if 'transition_table' in m:
        assert np.array_equal(orc.derive_reachable_states(m['transition_table']), m['reach'])
        assert np.array_equal(orc.reachable_probabilities_from_table(m['transition_table'], m['reach']), m['probs'])
    if 'reward_table' in m:
        S, A, R = m['reach'].shape
        O = m['rto'].shape[2]
        rr = m['reward_table'][np.arange(S)[:, None, None, None], np.arange(A)[None, :, None, None], m['reach'][:, :, :, None], np.arange(O)[None, None, None, :]]
        assert np.array_equal(orc.expected_rewards(m['rto'], rr), m['rbar'])
    elif len(m['end_states']):
        rr = orc.end_state_reachable_rewards(m['reach'], m['rto'].shape[2], m['end_states'])
        assert np.array_equal(orc.expected_rewards(m['rto'], rr), m['rbar'])


@pytest.mark.parametrize('tag', MODELS)
def test_backup_rows(tag):
    m, g = _model(tag), load_golden('backup_' + tag)
    reach = m['reach'].astype(np.int64)
    out = orc.backup(reach, m['rto'], m['rbar'], float(m['gamma']), g['beliefs'], g['alphas'])
    R = reach.shape[2]
    # action parity wherever the reference's own value gap exceeds the tolerance
    vals = np.sort(out['values'], axis=1)
    gap = vals[:, -1] - vals[:, -2] if vals.shape[1] > 1 else np.full(vals.shape[0], np.inf)
    decided = gap > GAP_TOL * np.maximum(1.0, np.abs(vals[:, -1]))
    assert np.array_equal(out['a_star'][decided], g['ref_row_action'][decided])
    same_a = out['a_star'] == g['ref_row_action']
    if R == 1:
        assert np.array_equal(out['alpha'][same_a], g['ref_row_alpha'][same_a])       # bit-exact
        assert same_a.all()
    else:
        np.testing.assert_allclose(out['alpha'][same_a], g['ref_row_alpha'][same_a], rtol=1e-12, atol=1e-13)


@pytest.mark.parametrize('tag', MODELS)
def test_backup_value_function_dedup_and_union(tag):
    m, g = _model(tag), load_golden('backup_' + tag)
    reach = m['reach'].astype(np.int64)
    out = orc.backup(reach, m['rto'], m['rbar'], float(m['gamma']), g['beliefs'], g['alphas'])
    rows, acts, first = orc.dedup_rows(out['alpha'], out['a_star'])
    assert rows.shape == g['ref_vf_alpha'].shape
    assert np.array_equal(acts, g['ref_vf_action'])
    np.testing.assert_allclose(rows, g['ref_vf_alpha'], rtol=1e-12, atol=1e-13)
    urows, uacts = orc.extend_union(rows, acts, g['alphas'], g['alpha_actions'])
    assert urows.shape == g['ref_vf_append_alpha'].shape
    assert np.array_equal(uacts, g['ref_vf_append_action'])
    np.testing.assert_allclose(urows, g['ref_vf_append_alpha'], rtol=1e-12, atol=1e-13)
    if reach.shape[2] == 1:
        assert np.array_equal(rows, g['ref_vf_alpha']) and np.array_equal(urows, g['ref_vf_append_alpha'])
    # belief-dominance filter (strict >)
    keep = orc.belief_dominance_filter(g['beliefs'], out['alpha'], g['alphas'])
    drows, dacts, _ = orc.dedup_rows(out['alpha'][keep], out['a_star'][keep])
    assert drows.shape == g['ref_vf_dom_alpha'].shape and np.array_equal(dacts, g['ref_vf_dom_action'])


def test_dedup_semantics_first_position_last_action():
    x, y, z = np.array([1.0, 2.0]), np.array([3.0, 4.0]), np.array([-0.0, 0.0])
    rows = np.stack([x, y, x, z, np.array([0.0, 0.0])])
    r, a, first = orc.dedup_rows(rows, np.array([0, 1, 1, 0, 2]))
    assert first.tolist() == [0, 1, 3, 4] and a.tolist() == [1, 1, 0, 2]      # -0.0 and 0.0 differ bytewise
    r2, a2 = orc.extend_union(rows[[0, 1]], np.array([0, 1]), rows[[3, 0]], np.array([2, 2]))
    assert a2.tolist() == [2, 1, 2] and np.array_equal(r2, np.stack([x, y, z]))   # old action wins, new order first


@pytest.mark.parametrize('tag', [t for t in MODELS if t != 'synth300'])
def test_belief_update_and_change(tag):
    m, g = _model(tag), load_golden('misc_' + tag)
    reach = m['reach'].astype(np.int64)
    with np.errstate(all='ignore'):
        if 'pairs' in g:
            for i, b in enumerate(g['beliefs']):
                for j, (a, o) in enumerate(g['pairs']):
                    assert np.array_equal(orc.belief_update(reach, m['rto'], b, a, o), g['ref_updates'][i, j], equal_nan=True)
        else:
            succ = orc.all_successors(reach, m['rto'], g['beliefs'])
            assert np.array_equal(succ, g['ref_updates'], equal_nan=True)
    chg = orc.compute_change(g['change_beliefs'], g['change_alphas_a'], g['change_alphas_b'])
    assert chg == pytest.approx(float(g['ref_change']), rel=1e-12, abs=1e-15)


def test_pairwise_sum_matches_numpy():
    rng = np.random.default_rng(0)
    for n in [0, 1, 7, 8, 9, 127, 128, 129, 255, 1000, 22021, 4097]:
        a = rng.random(n) * 10.0 ** rng.integers(-8, 8, n)
        assert orc.numpy_pairwise_sum(a) == float(np.sum(a)), n


@pytest.mark.parametrize('tag', ['tiger', 'grid4x4', 'grid4x4_noloop', 'tigergrid', 'hallway'])
def test_value_iteration_small(tag):
    m, g = _model(tag), load_golden('misc_' + tag)
    alphas, actions, iters = orc.vi_solve(m['reach'], m['probs'], m['rbar'], float(m['gamma']), 1e-6)
    assert iters == int(g['ref_vi_iters'])
    assert np.array_equal(actions, g['ref_vi_action'])
    np.testing.assert_allclose(alphas, g['ref_vi_alpha'], rtol=1e-13, atol=1e-13)


def test_value_iteration_known_answer_olfactory_nowrap():
    """The only artefact the reference pins: ValueFunctions/20231113_182429_value_function.csv (460 iterations)."""
    from pomdp_pbvi_exploration_b200.recipes import olfactory_wrap_model
    kat = load_golden('olf_nowrap_vi_kat')
    model = olfactory_wrap_model(wrap=False)
    alphas, actions, iters = orc.vi_solve(model.reachable_states, model.reachable_probabilities, model.expected_rewards_table, 0.99, 1e-4)
    assert iters == int(kat['ref_vi_iters']) == 460
    assert np.array_equal(actions, kat['kat_action'])
    assert np.max(np.abs(alphas - kat['kat_alpha'])) < 1e-12


def test_sawtooth_and_expansion_scores_tiger():
    m, g = _model('tiger'), load_golden('tiger_extras')
    for q, want in zip(g['ub_queries'], g['ref_ub_eval']):
        assert orc.sawtooth_reference(g['ub_corner'], g['ub_beliefs'], g['ub_values'], q) == pytest.approx(want, rel=1e-13)
        assert orc.sawtooth_intended(g['ub_corner'], g['ub_beliefs'], g['ub_values'], q) == pytest.approx(want, rel=1e-13)
    reach = m['reach']
    B = g['beliefs']
    succ = orc.all_successors(reach, m['rto'], B)
    n = g['ref_expand_ssea'].shape[0]
    d = orc.ssea_min_distances(B, succ)
    pick = np.argsort(d, axis=None)[::-1][:n]
    bi, ai, oi = np.unravel_index(pick, d.shape)
    assert np.array_equal(succ[bi, ai, oi], g['ref_expand_ssea'])
    res, probs, eps = orc.ger_scores(B, succ, g['alphas'], m['rto'], float(m['gamma']), float(m['min_reward']), float(m['max_reward']))
    bs, as_ = np.unravel_index(np.argsort(res, axis=None)[::-1][:n], res.shape)
    os_ = np.argmax(probs[bs, as_] * eps[bs, as_], axis=1)
    assert np.array_equal(succ[bs, as_, os_], g['ref_expand_ger'])
