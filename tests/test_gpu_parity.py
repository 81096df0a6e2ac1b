"""
GPU parity tests: the CUDA engine, called through the C ABI (pomdp_pbvi_exploration_b200._native -> libpbvi_b200.so),
against the CPU oracle (oracle/pbvi_oracle.py) and the golden fixtures produced by the unmodified reference.

Contract (BASELINE.json north_star): bit-exact alpha-argmax and action indices wherever the value gap exceeds the
tolerance; alpha entries within 1e-9 relative (bit-exact for reachable_state_count == 1 models, whose rows are sums of
single products in the reference's own operation order).
"""
import numpy as np
import pytest

from conftest import load_golden
from oracle import pbvi_oracle as orc

pytestmark = pytest.mark.gpu

GAP_TOL = 1e-9
MODELS = ['tiger', 'grid4x4', 'grid4x4_noloop', 'tigergrid', 'hallway', 'cheese', 'grid4x3', 'cit', 'synth300', 'olfactory_wrap']


@pytest.fixture(scope='module')
def torch_cuda():
    import torch
    assert torch.cuda.is_available(), 'these tests need a CUDA device'
    return torch


_dev_cache = {}


def device_model(tag):
    from pomdp_pbvi_exploration_b200._native import DeviceModel
    if tag not in _dev_cache:
        m = load_golden('model_' + tag)
        reach = m['reach'].astype(np.int64)
        probs = m['probs'] if 'probs' in m else np.full(reach.shape, 1.0 / reach.shape[2])
        _dev_cache[tag] = (DeviceModel(reach, probs, m['rto'], m['rbar']), m, reach, probs)
    return _dev_cache[tag]


def check_backup_against_oracle(dev, reach, rto, rbar, gamma, B, V, exact_rows):
    """Runs the engine and the oracle on the same inputs and applies the parity contract.  Returns the engine outputs."""
    alpha, act, vstar, value = [t.cpu().numpy() for t in dev.backup(B, V, gamma)]
    ref = orc.backup_chunked(reach, rto, rbar, gamma, B, V, chunk=64) if B.shape[0] > 64 else orc.backup(reach, rto, rbar, gamma, B, V)
    sc = np.concatenate([orc.backup(reach, rto, rbar, gamma, B[i:i + 64], V, return_scores=True)['scores'] for i in range(0, B.shape[0], 64)])
    # ---- v*: identical wherever the oracle's top-2 gap exceeds the tolerance; otherwise a near-maximal column
    top = np.sort(sc, axis=3)
    best = top[..., -1]
    gap = best - top[..., -2] if sc.shape[3] > 1 else np.full(best.shape, np.inf)
    scale = np.maximum(1.0, np.abs(best))
    decided = gap > GAP_TOL * scale
    assert np.array_equal(vstar[decided], ref['v_star'][decided])
    picked = np.take_along_axis(sc, vstar[..., None].astype(np.int64), axis=3)[..., 0]
    assert np.all(picked >= best - GAP_TOL * scale)
    # exact ties of a whole row (the all-zero (b,a,o) rows of sparse models): lowest index wins, as np.argmax
    flat = np.all(sc == sc[..., :1], axis=3)
    assert np.all(vstar[flat] == 0)
    # ---- values and a*
    np.testing.assert_allclose(value, ref['values'], rtol=1e-9, atol=1e-12)
    vs = np.sort(ref['values'], axis=1)
    agap = vs[:, -1] - vs[:, -2] if vs.shape[1] > 1 else np.full(vs.shape[0], np.inf)
    adecided = agap > GAP_TOL * np.maximum(1.0, np.abs(vs[:, -1]))
    assert np.array_equal(act[adecided], ref['a_star'][adecided])
    # a* must be the FIRST maximiser of the engine's own values (np.argmax semantics)
    assert np.array_equal(act, np.argmax(value, axis=1))
    # ---- alpha rows: compare where both engines selected the same (a*, v*[a*]) tuple
    same = (act == ref['a_star'])
    ours_sel = np.take_along_axis(vstar, act[:, None, None].astype(np.int64), axis=1)[:, 0, :]
    ref_sel = np.take_along_axis(ref['v_star'], ref['a_star'][:, None, None], axis=1)[:, 0, :]
    same &= np.all(ours_sel == ref_sel, axis=1)
    assert same.mean() > 0.5
    if exact_rows:
        assert np.array_equal(alpha[same], ref['alpha'][same])
    else:
        np.testing.assert_allclose(alpha[same], ref['alpha'][same], rtol=1e-9, atol=1e-12)
    return alpha, act, vstar, value, ref


@pytest.mark.parametrize('tag', MODELS)
def test_backup_golden(torch_cuda, tag):
    dev, m, reach, _ = device_model(tag)
    g = load_golden('backup_' + tag)
    gamma = float(m['gamma'])
    exact = reach.shape[2] == 1
    alpha, act, vstar, value, ref = check_backup_against_oracle(dev, reach, m['rto'], m['rbar'], gamma, g['beliefs'], g['alphas'], exact)
    # against the reference's own per-belief outputs
    vals = np.sort(ref['values'], axis=1)
    gap = vals[:, -1] - vals[:, -2]
    decided = gap > GAP_TOL * np.maximum(1.0, np.abs(vals[:, -1]))
    assert np.array_equal(act[decided], g['ref_row_action'][decided])
    same = act == g['ref_row_action']
    if exact:
        assert same.all() and np.array_equal(alpha, g['ref_row_alpha'])
    else:
        np.testing.assert_allclose(alpha[same], g['ref_row_alpha'][same], rtol=1e-9, atol=1e-12)
    # host-buffer entry point gives the same bytes as the device entry point
    ha, hact = dev.backup_host(g['beliefs'], g['alphas'], gamma)
    assert np.array_equal(ha, alpha) and np.array_equal(hact, act)
    # select + assemble of the selected tuples == fused call
    sel = np.take_along_axis(vstar, act[:, None, None].astype(np.int64), axis=1)[:, 0, :]
    rows = dev.backup_assemble(g['alphas'], gamma, act, sel).cpu().numpy()
    assert np.array_equal(rows, alpha)
    # keys accumulated while the rows are written == keys of a separate pass over the finished rows
    rows_t, keys = dev.backup_assemble(g['alphas'], gamma, act, sel, with_hash=True)
    assert torch_cuda.equal(keys, dev.row_hash(rows_t)) and np.array_equal(rows_t.cpu().numpy(), alpha)


@pytest.mark.parametrize('pinned', [False, True])
def test_backup_host_chunk_pipeline_equals_device_backup(torch_cuda, pinned):
    """`pbvi_backup_host` walks the beliefs in chunks of 2048 rows through a two-deep upload | kernels | download pipeline: 5000 beliefs
    = three chunks (the third re-uses the staging of the first), pageable and pinned host buffers, two calls in a row on one handle;
    alpha rows and actions equal the device entry point's, byte for byte."""
    import ctypes
    from pomdp_pbvi_exploration_b200.recipes import synthetic_sparse_model
    torch = torch_cuda
    model = synthetic_sparse_model(700, 4, 3, 1, seed=11)
    dev = model.device
    rng = np.random.default_rng(5)
    B = _sparse_beliefs(rng, 5000, 700, (3, 40, 200))
    V = rng.random((70, 700)) * (rng.random((70, 700)) < 0.4)
    want_rows, want_act, _, _ = dev.backup(B, V, 0.95)
    want_rows, want_act = want_rows.cpu().numpy(), want_act.cpu().numpy()
    hb, hv = torch.as_tensor(B), torch.as_tensor(V)
    out, act = torch.empty((5000, 700), dtype=torch.float64), torch.empty((5000,), dtype=torch.int32)
    if pinned:
        hb, hv, out, act = hb.pin_memory(), hv.pin_memory(), out.pin_memory(), act.pin_memory()
    for rep in range(2):
        out.fill_(-1.0)
        rc = dev._lib.pbvi_backup_host(dev._h, hb.data_ptr(), 5000, hv.data_ptr(), 70, ctypes.c_double(0.95), out.data_ptr(), act.data_ptr(),
                                       torch.cuda.current_stream().cuda_stream)
        assert rc == 0
        assert np.array_equal(out.numpy(), want_rows) and np.array_equal(act.numpy(), want_act)
    # the arena is handed back to the next call intact: a plain device backup right after gives the same rows again
    again, _, _, _ = dev.backup(B[:300], V, 0.95)
    assert np.array_equal(again.cpu().numpy(), want_rows[:300])


@pytest.mark.parametrize('case', ['synthetic', 'synthetic_dense', 'tiger', 'olfactory_wrap'])
def test_backup_host_unique_equals_the_solver_backup(torch_cuda, case):
    """`pbvi_backup_host_unique` -- the reference's whole backup from host buffers in ONE library call, ValueFunction-constructor dedup
    included -- returns the rows and actions of `PBVI_Solver.backup(append=False, belief_dominance_prune=False)`, in the same order,
    byte for byte.  Sparse sets of 2048 rows or more travel packed (host threads + unpack kernel: `synthetic`, `olfactory_wrap`), dense
    or small ones through the chunked dense pipeline (`synthetic_dense`: three chunks; tiger: R = 2, many tuples per row)."""
    from pomdp_pbvi_exploration_b200 import BeliefSet, PBVI_Solver, ValueFunction
    from pomdp_pbvi_exploration_b200.recipes import olfactory_wrap_model, perseus_walk_beliefs, synthetic_sparse_model
    rng = np.random.default_rng(21)
    if case == 'synthetic':
        model, gamma = synthetic_sparse_model(700, 4, 3, 1, seed=11), 0.95
        B = _sparse_beliefs(rng, 5000, 700, (1, 2, 3, 40))
        B[100:140] = B[0]                                   # repeated beliefs: repeated tuples
        V, acts = rng.random((70, 700)) * (rng.random((70, 700)) < 0.4), rng.integers(0, 4, 70)
    elif case == 'synthetic_dense':
        model, gamma = synthetic_sparse_model(700, 4, 3, 1, seed=11), 0.95
        B = rng.dirichlet(np.ones(700), size=4300)
        B[7] = B[4200]
        V, acts = rng.random((70, 700)), rng.integers(0, 4, 70)
    elif case == 'tiger':
        g = load_golden('backup_tiger')
        dev, mm, _, _ = device_model('tiger')
        B, V, gamma = g['beliefs'], g['alphas'], float(mm['gamma'])
        rows, acts_out = dev.backup_host_unique(B, V, gamma)
        alpha, act, _, _ = dev.backup(B, V, gamma)
        alpha, act = alpha.cpu().numpy(), act.cpu().numpy()
        table = {}
        for i in range(alpha.shape[0]):                     # the reference's dict comprehension: first position, last action
            k = alpha[i].tobytes()
            table[k] = (table[k][0] if k in table else i, act[i])
        want_rows = np.stack([alpha[p] for p, _ in table.values()])
        want_acts = np.array([a for _, a in table.values()])
        assert np.array_equal(rows, want_rows) and np.array_equal(acts_out, want_acts)
        return
    else:
        model, gamma = olfactory_wrap_model(), 0.99
        g = load_golden('backup_olfactory_wrap')
        B = perseus_walk_beliefs(model, 2300, seed=2)
        V, acts = g['alphas'], g['alpha_actions']
    vf = ValueFunction(model, V, acts)
    solver = PBVI_Solver(gamma=gamma, eps=1e-6, expand_function='perseus')
    want = solver.backup(model, BeliefSet(model, B), vf, append=False, belief_dominance_prune=False)
    want_rows, want_acts = want.numpy()
    rows, acts_out = model.device.backup_host_unique(B, vf.alpha_vector_array.cpu().numpy(), gamma)
    assert rows.shape == want_rows.shape and np.array_equal(rows, want_rows) and np.array_equal(acts_out, want_acts)
    # too little room: the count comes back with the error
    import ctypes
    n = ctypes.c_int()
    small = np.empty((1, model.state_count))
    a1 = np.empty(1, dtype=np.int32)
    Bc, Vc = np.ascontiguousarray(B), np.ascontiguousarray(vf.alpha_vector_array.cpu().numpy())
    rc = model.device._lib.pbvi_backup_host_unique(model.device._h, Bc.ctypes.data, Bc.shape[0], Vc.ctypes.data, Vc.shape[0], gamma,
                                                   small.ctypes.data, 1, a1.ctypes.data, ctypes.byref(n), None)
    assert (rc != 0) == (want_rows.shape[0] > 1) and n.value == want_rows.shape[0]


def _sparse_beliefs(rng, n, S, ks):
    B = np.zeros((n, S))
    for i in range(n):
        k = min(S, ks[i % len(ks)])
        idx = rng.choice(S, k, replace=False)
        B[i, idx] = rng.dirichlet(np.ones(k))
    return B


@pytest.mark.parametrize('S,A,O,R,nB,nV', [
    (37, 2, 2, 1, 1, 1), (100, 3, 2, 1, 127, 5), (300, 4, 3, 1, 129, 129), (515, 3, 4, 1, 300, 257),
    (64, 3, 3, 2, 40, 17), (200, 4, 2, 5, 131, 130), (1000, 2, 2, 3, 64, 260),
])
def test_backup_synthetic_shapes(torch_cuda, S, A, O, R, nB, nV):
    """Tile-boundary coverage (ragged M / N / K tiles, one belief, one alpha) on the configs[4] recipe."""
    from pomdp_pbvi_exploration_b200._native import DeviceModel
    rng = np.random.default_rng(S * 7 + R)
    reach = rng.integers(0, S, (S, A, R))
    probs = np.full(reach.shape, 1.0 / R)
    obs = rng.random((S, A, O)); obs /= obs.sum(2, keepdims=True)
    obs[rng.random((S, A, O)) < 0.3] = 0.0            # impossible observations -> exact-zero score rows
    rto = orc.build_rto(reach, probs, obs)
    rbar = orc.expected_rewards(rto, orc.end_state_reachable_rewards(reach, O, [S // 2]))
    B = _sparse_beliefs(rng, nB, S, [S, 40, 5, 1])
    V = rng.random((nV, S))
    dev = DeviceModel(reach, probs, rto, rbar)
    check_backup_against_oracle(dev, reach, rto, rbar, 0.95, B, V, exact_rows=(R == 1))
    dev.close()


def test_backup_all_zero_and_duplicate_rows(torch_cuda):
    """Empty support intersections: beliefs that only see impossible observations give all-zero score rows -> v* = 0."""
    from pomdp_pbvi_exploration_b200._native import DeviceModel
    S, A, O = 50, 2, 3
    reach = ((np.arange(S)[:, None, None] + np.arange(1, A + 1)[None, :, None]) % S).astype(np.int64)
    obs = np.zeros((S, A, O)); obs[:, :, 0] = 1.0                                  # only o = 0 ever happens
    rto = orc.build_rto(reach, np.ones(reach.shape), obs)
    rbar = np.zeros((S, A)); rbar[:, 1] = 0.0
    dev = DeviceModel(reach, None, rto, rbar)
    rng = np.random.default_rng(3)
    B = _sparse_beliefs(rng, 10, S, [3])
    V = rng.random((9, S))
    alpha, act, vstar, value = [t.cpu().numpy() for t in dev.backup(B, V, 0.9)]
    assert np.all(vstar[:, :, 1:] == 0)
    ref = orc.backup(reach, rto, rbar, 0.9, B, V)
    assert np.array_equal(vstar, ref['v_star']) and np.array_equal(act, ref['a_star']) and np.array_equal(alpha, ref['alpha'])
    dev.close()


@pytest.mark.parametrize('tag', ['tiger', 'hallway', 'olfactory_wrap'])
def test_max_values_and_change(torch_cuda, tag):
    dev, m, reach, _ = device_model(tag)
    g = load_golden('misc_' + tag)
    B = g['change_beliefs']
    for V in (g['change_alphas_a'], g['change_alphas_b']):
        mx, arg = [t.cpu().numpy() for t in dev.max_values(B, V)]
        rmx, rarg = orc.max_values(B, V)
        np.testing.assert_allclose(mx, rmx, rtol=1e-12, atol=1e-14)
        prod = B @ V.T
        top = np.sort(prod, axis=1)
        decided = (top[:, -1] - top[:, -2] > GAP_TOL * np.maximum(1, np.abs(top[:, -1]))) if V.shape[0] > 1 else np.ones(len(B), bool)
        assert np.array_equal(arg[decided], rarg[decided])
    va = dev.max_values(B, g['change_alphas_a'])[0]
    vb = dev.max_values(B, g['change_alphas_b'])[0]
    change = float((vb - va).abs().max())
    assert change == pytest.approx(float(g['ref_change']), rel=1e-9, abs=1e-12)


@pytest.mark.parametrize('tag', [t for t in MODELS if t != 'synth300'])
def test_belief_update_bit_exact(torch_cuda, tag):
    """Belief.update: bincount accumulation order + NumPy pairwise normaliser reproduced bit for bit (NaN rows included)."""
    dev, m, reach, _ = device_model(tag)
    g = load_golden('misc_' + tag)
    B = g['beliefs']
    if 'pairs' in g:
        pairs = g['pairs']
        bb = np.repeat(B, len(pairs), axis=0)
        aa = np.tile(pairs[:, 0], len(B)); oo = np.tile(pairs[:, 1], len(B))
        out, mass = dev.belief_update(bb, aa, oo)
        got = out.cpu().numpy().reshape(len(B), len(pairs), -1)
        assert np.array_equal(got, g['ref_updates'], equal_nan=True)
        # one to four rows take the multi-block normaliser: same bytes, same masses
        for k in (1, 3):
            out_k, mass_k = dev.belief_update(bb[:k], aa[:k], oo[:k])
            assert np.array_equal(out_k.cpu().numpy(), out[:k].cpu().numpy(), equal_nan=True)
            assert np.array_equal(mass_k.cpu().numpy(), mass[:k].cpu().numpy(), equal_nan=True)
    else:
        succ, mass = dev.belief_successors(B)
        assert np.array_equal(succ.cpu().numpy(), g['ref_updates'], equal_nan=True)
        probs = dev.observation_probabilities(B).cpu().numpy()
        want = np.einsum('bs,saor->bao', B, m['rto'])
        np.testing.assert_allclose(probs, want, rtol=1e-12, atol=1e-15)
        np.testing.assert_allclose(mass.cpu().numpy(), want, rtol=1e-12, atol=1e-15)
    # un-normalised projection is the bincount itself
    a0 = np.zeros(len(B), dtype=np.int32); o0 = np.zeros(len(B), dtype=np.int32)
    raw, _ = dev.belief_update(B, a0, o0, normalise=False)
    want = orc.belief_update_batch(reach, m['rto'], B, a0, o0, normalise=False)
    assert np.array_equal(raw.cpu().numpy(), want)


def test_row_hash_and_equality(torch_cuda):
    dev, m, reach, _ = device_model('hallway')
    rng = np.random.default_rng(1)
    S = dev.S
    rows = rng.random((40, S))
    rows[7] = rows[3]; rows[21] = rows[3]
    rows[9, 5] = 0.0; rows[10] = rows[9]; rows[10, 5] = -0.0          # -0.0 and 0.0 differ bytewise
    h = dev.row_hash(rows).cpu().numpy()
    assert np.array_equal(h[7], h[3]) and np.array_equal(h[21], h[3])
    assert not np.array_equal(h[9], h[10])
    keys = {tuple(x) for x in h.tolist()}
    assert len(keys) == 38
    swapped = rows[0].copy(); swapped[[1, 2]] = swapped[[2, 1]]         # position dependence
    assert not np.array_equal(dev.row_hash(swapped[None]).cpu().numpy()[0], h[0])
    flags = dev.rows_equal(rows, [3, 3, 9, 0], rows, [7, 21, 10, 1]).cpu().numpy()
    assert flags.tolist() == [1, 1, 0, 0]


@pytest.mark.parametrize('tag', ['tiger', 'grid4x4', 'tigergrid', 'hallway'])
def test_vi_sweep_and_solve(torch_cuda, tag):
    dev, m, reach, probs = device_model(tag)
    g = load_golden('misc_' + tag)
    gamma = float(m['gamma'])
    v = np.max(m['rbar'], axis=1)
    alpha, vnew = dev.vi_sweep(v, gamma)
    want = orc.vi_sweep(reach, probs, m['rbar'], gamma, v)
    np.testing.assert_allclose(alpha.cpu().numpy(), want, rtol=1e-13, atol=1e-13)
    np.testing.assert_allclose(vnew.cpu().numpy(), want.max(0), rtol=1e-13, atol=1e-13)


def test_prune_sawtooth_distance(torch_cuda):
    dev, m, reach, _ = device_model('tiger')
    g = load_golden('tiger_extras')
    got = dev.sawtooth(g['ub_corner'], g['ub_beliefs'], g['ub_values'], g['ub_queries']).cpu().numpy()
    np.testing.assert_allclose(got, g['ref_ub_eval'], rtol=1e-12)
    assert dev.sawtooth(g['ub_corner'], np.zeros((0, 2)), np.zeros(0), g['ub_queries'][:1]).item() == pytest.approx(
        float(g['ub_queries'][0] @ g['ub_corner']))
    dev2, m2, reach2, _ = device_model('hallway')
    rng = np.random.default_rng(2)
    al = rng.random((30, dev2.S))
    al[4] = al[2] - 0.1; al[11] = al[7]; al[20] = np.minimum(al[1], al[3]) - 1e-3
    keep = dev2.prune_dominated(al).cpu().numpy().astype(bool)
    assert np.array_equal(np.flatnonzero(keep), orc.prune_pointwise_dominated(al))
    B = _sparse_beliefs(rng, 12, dev2.S, [dev2.S, 7])
    succ = orc.all_successors(reach2, m2['rto'], B[:3])
    with np.errstate(all='ignore'):
        want = orc.ssea_min_distances(B, succ)
    got = dev2.min_l2_distance(B, succ.reshape(-1, dev2.S)).cpu().numpy().reshape(want.shape)
    np.testing.assert_allclose(got, want, rtol=1e-10, atol=1e-12, equal_nan=True)


def test_error_convention(torch_cuda):
    from pomdp_pbvi_exploration_b200 import _native
    dev, m, reach, _ = device_model('tiger')
    with pytest.raises(ValueError):
        dev.backup(np.array([[0.5, 0.5]]), np.zeros((1, 2)), -1.0)
    with pytest.raises(ValueError):
        _native.DeviceModel(np.full((2, 1, 1), 5), None, np.ones((2, 1, 1, 1)), np.zeros((2, 1)))   # landing state out of range
    assert b'reachable_states' in _native.load_library().pbvi_last_error()


def test_olfactory_full_size_properties(torch_cuda):
    """
    BASELINE-size model (S = 22021): properties that do not need the CPU oracle on the whole input --
    determinism, tile independence (a row's result does not depend on its tile mates), alpha-order equivariance --
    plus the oracle on a sample of rows.
    """
    from pomdp_pbvi_exploration_b200.recipes import olfactory_wrap_model, perseus_walk_beliefs
    torch = torch_cuda
    model = olfactory_wrap_model()
    dev = model.device
    rng = np.random.default_rng(0)
    B = perseus_walk_beliefs(model, 700, seed=0)
    g = load_golden('backup_olfactory_wrap')
    V = np.concatenate([g['alphas'], rng.random((300 - g['alphas'].shape[0], dev.S)) * 0.05])
    alpha, act, vstar, value = dev.backup(B, V, 0.99)
    alpha2, act2, vstar2, value2 = dev.backup(B, V, 0.99)
    assert torch.equal(alpha, alpha2) and torch.equal(vstar, vstar2) and torch.equal(value, value2)
    idx = np.sort(rng.choice(700, 150, replace=False))
    a3, act3, vs3, val3 = dev.backup(B[idx], V, 0.99)
    tidx = torch.as_tensor(idx, device=alpha.device)
    assert torch.equal(vs3, vstar[tidx]) and torch.equal(act3, act[tidx]) and torch.equal(a3, alpha[tidx])
    # reversing the alpha order maps decided v* to V-1-v*
    a4, act4, vs4, val4 = dev.backup(B[idx], V[::-1].copy(), 0.99)
    torch.testing.assert_close(val4, val3, rtol=1e-12, atol=1e-14)
    # oracle on a sample of rows (full parity contract), and the sample's rows equal the full run's rows
    samp = idx[:24]
    tsamp = torch.as_tensor(samp, device=alpha.device)
    a5, act5, vs5, val5, ref = check_backup_against_oracle(dev, model.reachable_states, model.reachable_transitional_observation_table,
                                                           model.expected_rewards_table, 0.99, B[samp], V, exact_rows=True)
    assert np.array_equal(vs5, vstar[tsamp].cpu().numpy()) and np.array_equal(a5, alpha[tsamp].cpu().numpy())
    # reversing the alpha order maps every decided v* to V-1-v*
    sc = orc.backup(model.reachable_states, model.reachable_transitional_observation_table, model.expected_rewards_table, 0.99, B[samp], V,
                    return_scores=True)['scores']
    top = np.sort(sc, axis=3)
    decided = (top[..., -1] - top[..., -2]) > GAP_TOL * np.maximum(1.0, np.abs(top[..., -1]))
    rev = vs4[:24].cpu().numpy()
    assert np.array_equal((V.shape[0] - 1 - rev)[decided], ref['v_star'][decided])


@pytest.mark.parametrize('negative', [False, True])
def test_backup_sparse_alphas_and_sign_flags(torch_cuda, negative):
    """
    Alpha-side zero skipping and the exact-zero shortcut of the value pass: alpha vectors that vanish on most states (whole
    256-wide alpha tiles are zero on most chunks), beliefs that miss their support entirely (all values exactly 0), and --
    with `negative` -- alphas of mixed sign, for which the shortcut must stay off.  Parity contract as everywhere else.
    """
    from pomdp_pbvi_exploration_b200._native import DeviceModel
    rng = np.random.default_rng(17 + int(negative))
    S, A, O, nB, nV = 700, 3, 3, 150, 530
    reach = rng.integers(0, S, (S, A, 1))
    obs = rng.random((S, A, O)); obs /= obs.sum(2, keepdims=True)
    obs[rng.random((S, A, O)) < 0.25] = 0.0
    rto = orc.build_rto(reach, np.ones(reach.shape), obs)
    rbar = orc.expected_rewards(rto, orc.end_state_reachable_rewards(reach, O, [S // 2]))
    V = np.zeros((nV, S))
    for v in range(nV):                                           # support: a window of states that depends on the alpha TILE
        lo = (v // 256) * 200 + rng.integers(0, 40)
        V[v, lo:lo + 60] = rng.random(60)
    if negative:
        V[rng.random(V.shape) < 0.02] *= -1.0
    B = _sparse_beliefs(rng, nB, S, [S, 40, 5, 1])
    B[::7] = 0.0
    B[::7, 650:660] = 0.1                                         # beliefs far from every alpha support -> all-zero scores
    dev = DeviceModel(reach, None, rto, rbar)
    alpha, act, vstar, value, ref = check_backup_against_oracle(dev, reach, rto, rbar, 0.95, B, V, exact_rows=True)
    far = np.arange(0, nB, 7)
    if not negative:
        assert np.all(value[far][ref['values'][far] == 0.0] == 0.0)
    mx, arg = [t.cpu().numpy() for t in dev.max_values(B, V)]
    rmx, rarg = orc.max_values(B, V)
    np.testing.assert_allclose(mx, rmx, rtol=1e-12, atol=1e-15)
    prod = np.sort(B @ V.T, axis=1)
    decided = prod[:, -1] - prod[:, -2] > GAP_TOL * np.maximum(1, np.abs(prod[:, -1]))
    assert np.array_equal(arg[decided], rarg[decided])
    dev.close()


@pytest.mark.parametrize('n,pool,width,dtype', [(1, 1, 4, 'int32'), (7, 3, 1, 'int32'), (1000, 37, 4, 'int32'), (5000, 5000, 2, 'int64'),
                                                (20000, 900, 5, 'int32'), (3001, 2, 2, 'int64')])
def test_group_keys_device_equals_dict_insertion(torch_cuda, n, pool, width, dtype):
    """`pbvi_group_keys` == the reference's `{row.tobytes(): value}` dict (src/mdp.py:668-669): groups in order of first
    occurrence, first position, last occurrence (or the record of largest rank)."""
    torch = torch_cuda
    dev, m, reach, _ = device_model('tiger')
    rng = np.random.default_rng(n + width)
    hi = 2 ** 31 - 1 if dtype == 'int32' else 2 ** 62
    base = rng.integers(-hi, hi, (pool, width)).astype(dtype)
    keys = base[rng.integers(0, pool, n)]
    table = {}
    for i, k in enumerate(keys):
        kb = k.tobytes()
        if kb in table:
            table[kb][1] = i
        else:
            table[kb] = [i, i, len(table)]
    want_first = np.array([v[0] for v in table.values()])
    want_last = np.array([v[1] for v in table.values()])
    want_inv = np.array([table[k.tobytes()][2] for k in keys])
    first, last, inv = dev.group_keys(torch.as_tensor(keys).cuda(), want_inverse=True)
    assert np.array_equal(first.cpu().numpy(), want_first) and np.array_equal(last.cpu().numpy(), want_last)
    assert np.array_equal(inv.cpu().numpy(), want_inv)
    # with ranks: `last` is the record of largest rank inside the group
    rank = rng.permutation(n).astype(np.int32)
    first2, owner, _ = dev.group_keys(torch.as_tensor(keys).cuda(), rank=torch.as_tensor(rank).cuda())
    want_owner = np.array([np.flatnonzero(want_inv == g)[np.argmax(rank[want_inv == g])] for g in range(len(table))]) if n <= 5000 else None
    assert np.array_equal(first2.cpu().numpy(), want_first)
    if want_owner is not None:
        assert np.array_equal(owner.cpu().numpy(), want_owner)


def test_confirm_groups_and_backup_dedup(torch_cuda):
    """Byte confirmation of key groups, and PBVI_Solver.rows_from_tuples on tuples that generate identical rows: first position,
    action of the tuple whose last belief comes latest (reference ValueFunction ctor semantics)."""
    torch = torch_cuda
    dev, m, reach, _ = device_model('hallway')
    rng = np.random.default_rng(5)
    rows = rng.random((30, dev.S))
    rows[11] = rows[4]; rows[29] = rows[4]; rows[17] = rows[16]
    t = torch.as_tensor(rows).cuda()
    keys = dev.row_hash(t)
    first, last, inv = dev.group_keys(keys, want_inverse=True)
    assert first.shape[0] == 27 and dev.confirm_groups(t, first, inv)
    bad = keys.clone()
    bad[20] = bad[2]                                            # a forged key match between different rows
    first_b, _, inv_b = dev.group_keys(bad, want_inverse=True)
    assert first_b.shape[0] == 26 and not dev.confirm_groups(t, first_b, inv_b)


@pytest.mark.parametrize('interleaved', [False, True])
@pytest.mark.parametrize('world,block_rows,width', [(1, 5, 3), (2, 40, 4), (8, 257, 4), (3, 9, 1)])
def test_group_record_blocks_equals_host_merge(torch_cuda, world, block_rows, width, interleaved):
    """`pbvi_group_record_blocks` (merge of the all-gathered tuple blocks of a sharded backup) == the host twin used by the gloo
    tests: ragged record counts per rank (including an empty and an overflowing block), duplicate tuples across ranks."""
    from pomdp_pbvi_exploration_b200.parallel import merge_blocks_host
    torch = torch_cuda
    dev, m, reach, _ = device_model('tiger')
    rng = np.random.default_rng(world * 100 + block_rows)
    pool = rng.integers(0, 50, (max(3, block_rows // 2), width))
    blocks = rng.integers(0, 1000, (world, block_rows, width + 2)).astype(np.int32)     # padding rows hold garbage
    pos = 0
    for r in range(world):
        u = int(rng.integers(0, block_rows)) if r else block_rows - 1
        if r == 1:
            u = 0
        blocks[r, 0, :] = u
        pick = rng.choice(pool.shape[0], min(u, pool.shape[0]), replace=False)
        u = pick.shape[0]
        blocks[r, 0, :] = u
        blocks[r, 1:1 + u, :width] = pool[pick]
        firsts = pos + np.sort(rng.choice(10 * block_rows, u, replace=False))
        blocks[r, 1:1 + u, width] = firsts
        blocks[r, 1:1 + u, width + 1] = firsts + rng.integers(0, 5, u)
        pos += 10 * block_rows + 5
    if interleaved:
        # the sharded solve's ownership: positions of different ranks interleave (and do not follow the buffer order); the merge
        # must order by position, not by where a record sits in the gathered buffer
        valid = np.concatenate([r * block_rows + 1 + np.arange(int(blocks[r, 0, 0])) for r in range(world)]).astype(np.int64)
        flat_v = blocks.reshape(world * block_rows, width + 2)
        perm = rng.permutation(valid.shape[0])
        flat_v[valid, width] = np.sort(flat_v[valid, width])[perm]
        flat_v[valid, width + 1] = flat_v[valid, width] + rng.integers(0, 5, valid.shape[0])
    flat = torch.as_tensor(blocks.reshape(world * block_rows, width + 2))
    want_first, want_last, want_max = merge_blocks_host(flat, world, block_rows, width)
    first, last, mx = dev.group_record_blocks(flat.cuda(), world, block_rows, width)
    assert mx == want_max
    assert np.array_equal(first.cpu().numpy(), want_first.numpy()) and np.array_equal(last.cpu().numpy(), want_last.numpy())
    # an overflowing header is reported, not silently truncated
    blocks[0, 0, :] = block_rows + 3
    _, _, mx2 = dev.group_record_blocks(torch.as_tensor(blocks.reshape(world * block_rows, width + 2)).cuda(), world, block_rows, width)
    assert mx2 == block_rows + 3


@pytest.mark.parametrize('n,L', [(1, 1), (5, 3), (300, 130), (257, 22021), (64, 4096)])
def test_pack_unpack_rows_bytewise(torch_cuda, n, L):
    """Sparse-row transport: host pack (pbvi_pack_rows_host) + device unpack (pbvi_unpack_rows) reproduces the rows byte for byte,
    including -0.0, NaN payloads, all-zero rows, dense rows and a ragged last chunk."""
    torch = torch_cuda
    dev, m, reach, _ = device_model('tiger')
    rng = np.random.default_rng(n * 31 + L)
    rows = np.zeros((n, L))
    for i in range(n):
        k = int(rng.integers(0, L + 1))
        if i % 3 == 0:
            lo = int(rng.integers(0, L))
            rows[i, lo:lo + k] = rng.random(min(k, L - lo))
        elif i % 3 == 1:
            rows[i, rng.choice(L, min(L, 7), replace=False)] = rng.random(min(L, 7))
    rows[0, -1] = -0.0
    if n > 2:
        rows[2] = rng.random(L)                                    # dense row
        rows[1, 0] = np.nan
    t = torch.from_numpy(rows)
    n_c, W = dev.pack_geometry(L)
    bm = torch.empty((n, W), dtype=torch.int32)
    rs = torch.empty((n + 1,), dtype=torch.int32)
    pk = torch.empty((n * n_c * 4 + 4,), dtype=torch.float64)
    total = dev.pack_rows_host(t, bm, rs, pk)
    assert total == int(rs[n]) and total <= n * n_c
    out = torch.full((n, L), 7.0, dtype=torch.float64, device='cuda')
    dev.unpack_rows(bm.cuda(), rs.cuda(), pk[:max(4, total * 4)].cuda(), out)
    assert np.array_equal(out.cpu().numpy().view(np.uint64), rows.view(np.uint64))
    # the same rows packed in slabs of 7 (regions at a fixed stride), unpacked by one launch
    SL = 7
    n_slabs = -(-n // SL)
    region = SL * n_c * 4 + 4
    bm2 = torch.empty((n, W), dtype=torch.int32)
    rs2 = torch.zeros((n_slabs, SL + 1), dtype=torch.int32)
    pk2 = torch.zeros((n_slabs * region,), dtype=torch.float64)
    for i in range(n_slabs):
        r0, r1 = i * SL, min(n, (i + 1) * SL)
        dev.pack_rows_host(t[r0:r1], bm2[r0:r1], rs2[i], pk2[i * region:(i + 1) * region])
    out2 = torch.full((n, L), 7.0, dtype=torch.float64, device='cuda')
    dev.unpack_rows(bm2.cuda(), rs2.cuda(), pk2.cuda(), out2, slab_rows=SL, region_chunks=region // 4)
    assert np.array_equal(out2.cpu().numpy().view(np.uint64), rows.view(np.uint64))


@pytest.mark.parametrize('S,n_ub,n_q', [(60, 5, 3), (9000, 40, 18), (20011, 33, 7)])
def test_sawtooth_sparse_stored_beliefs(torch_cuda, S, n_ub, n_q):
    """HSVI upper bound on sparse stored beliefs (support compacted per 8192-state range in shared memory, one block per stored
    belief): equals the oracle's support-restricted sawtooth; a query equal to a stored belief, a stored belief whose support
    misses the query (ratio 0), supports spanning several ranges."""
    from pomdp_pbvi_exploration_b200._native import DeviceModel
    rng = np.random.default_rng(S + n_ub)
    reach = rng.integers(0, S, (S, 2, 1))
    obs = np.full((S, 2, 2), 0.5)
    rto = orc.build_rto(reach, np.ones(reach.shape), obs)
    dev = DeviceModel(reach, None, rto, np.zeros((S, 2)))
    corner = rng.random(S) * 10
    ub = _sparse_beliefs(rng, n_ub, S, [S, min(S, 300), 17, 1])
    queries = _sparse_beliefs(rng, n_q, S, [S, min(S, 500), 40])
    queries[0] = ub[1]
    ub_v = ub @ corner - rng.random(n_ub)                         # stored values below the corner interpolation
    got = dev.sawtooth(corner, ub, ub_v, queries).cpu().numpy()
    want = np.array([orc.sawtooth_intended(corner, ub, ub_v, q) for q in queries])
    np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-13)
    dev.close()


@pytest.mark.parametrize('tag', ['tiger', 'hallway', 'olfactory_wrap'])
def test_belief_trajectory_equals_chained_updates(torch_cuda, tag):
    """pbvi_belief_trajectory (the FSVI chain: project + multi-block pairwise normaliser per step, restarts from b0) is bit-identical
    to chaining pbvi_belief_update one step at a time."""
    torch = torch_cuda
    dev, m, reach, _ = device_model(tag)
    g = load_golden('backup_' + tag)
    rng = np.random.default_rng(11)
    b0 = torch.as_tensor(g['beliefs'][0]).cuda()
    n = 25
    actions, observations, resets = [], [], []
    cur = b0
    want = []
    for i in range(n):
        a = int(rng.integers(dev.A))
        probs = dev.observation_probabilities(cur[None, :])[0, a].cpu().numpy()
        o = int(rng.choice(np.flatnonzero(probs > 0)))
        nxt, _ = dev.belief_update(cur[None, :], [a], [o])
        want.append(nxt[0])
        reset = i % 9 == 8
        actions.append(a); observations.append(o); resets.append(reset)
        cur = b0 if reset else nxt[0]
    chain = dev.belief_trajectory(b0, actions, observations, resets)
    assert torch.equal(chain, torch.stack(want))


@pytest.mark.parametrize('tag', ['tiger', 'grid4x4', 'olfactory_wrap'])
def test_perseus_walk_equals_host_driven_steps(torch_cuda, tag):
    """`pbvi_perseus_walk` (observations drawn on the device from host uniforms) == the step-by-step form of the reference's
    expand_perseus (src/pomdp.py:2040-2054): P(o|b,a) read back, `np.random.choice(observations, p=...)` on the host, one update
    -- same beliefs bit for bit, same observations, same state of the legacy RNG stream afterwards."""
    dev, m, reach, _ = device_model(tag)
    A, O = m['rto'].shape[1], m['rto'].shape[2]
    n = 40
    b0 = m['start']
    np.random.seed(21)
    cur = torch_cuda.as_tensor(b0).cuda()
    want_rows, want_obs = [], []
    for _ in range(n):
        a = int(np.random.choice(np.arange(A), size=1)[0])
        p = dev.observation_probabilities(cur[None, :])[0, a].cpu().numpy()
        o = int(np.random.choice(np.arange(O), size=1, p=p)[0])
        cur = dev.belief_update(cur[None, :], [a], [o])[0][0]
        want_rows.append(cur.cpu().numpy())
        want_obs.append(o)
    tail_want = np.random.random_sample()
    np.random.seed(21)
    acts, us = np.empty(n, dtype=np.int32), np.empty(n)
    for i in range(n):
        acts[i] = int(np.random.choice(np.arange(A), size=1)[0])
        us[i] = np.random.random_sample()
    tail_got = np.random.random_sample()
    rows, obs = dev.perseus_walk(b0, acts, us, want_observations=True)
    assert tail_got == tail_want
    assert np.array_equal(obs.cpu().numpy(), np.array(want_obs))
    assert np.array_equal(rows.cpu().numpy(), np.stack(want_rows), equal_nan=True)


@pytest.mark.parametrize('tag,n_b,n_src', [('hallway', 70, 9), ('grid4x4', 300, 40), ('olfactory_wrap', 90, 6), ('tiger', 5, 5)])
def test_min_l2_and_batched_successors_at_scale(torch_cuda, tag, n_b, n_src):
    """SSEA ingredients beyond tiger (the reference's SSEA crashes there, so the ingredients are compared): all successors of
    a belief chunk from ONE launch == per-belief updates of the oracle (bit-exact, NaN rows included), and the tiled distance kernel ==
    the reference's diff / einsum / sqrt / min on the possible successors (tile edges: n_b, n_src, S not multiples of the tiles)."""
    dev, m, reach, _ = device_model(tag)
    g = load_golden('backup_' + tag)
    rng = np.random.default_rng(5)
    S = dev.S
    B = g['beliefs']
    if B.shape[0] < n_b:
        extra = _sparse_beliefs(rng, n_b - B.shape[0], S, [S, max(1, S // 3), 2])
        B = np.concatenate([B, extra])
    B = np.ascontiguousarray(B[:n_b])
    src = B[:n_src]
    succ, mass = dev.belief_successors(src)
    with np.errstate(all='ignore'):
        want_succ = orc.all_successors(reach, m['rto'], src)
    assert np.array_equal(succ.cpu().numpy(), want_succ, equal_nan=True)
    possible = mass.cpu().numpy().reshape(-1) > 0
    assert np.array_equal(possible, ~np.isnan(want_succ.reshape(-1, S)).any(axis=1))
    cand = succ.reshape(-1, S)
    got = dev.min_l2_distance(B, cand).cpu().numpy()
    with np.errstate(all='ignore'):
        want = orc.ssea_min_distances(B, want_succ).reshape(-1)
    np.testing.assert_allclose(got, want, rtol=1e-10, atol=1e-13, equal_nan=True)
    assert np.isnan(got[~possible]).all() and np.isfinite(got[possible]).all()
    # successors that are already in the set are at distance exactly 0, as in the reference
    in_set = dev.min_l2_distance(np.concatenate([B, want_succ.reshape(-1, S)[possible][:3]]), cand[torch_cuda.as_tensor(np.flatnonzero(possible)[:3]).cuda()])
    assert np.all(in_set.cpu().numpy() == 0.0)


def test_calls_on_two_streams_of_one_handle_are_ordered(torch_cuda):
    """A handle's scratch (arena, sign flags, tile counter) is reused in stream order; a call that arrives on another stream than its
    predecessor waits on the device for the predecessor's work (pbvi_model::last_stream), so alternating streams gives the results of a
    single stream."""
    torch = torch_cuda
    dev, m, reach, _ = device_model('olfactory_wrap')
    g = load_golden('backup_olfactory_wrap')
    B = torch.as_tensor(g['beliefs']).cuda()
    V = torch.as_tensor(g['alphas']).cuda()
    gamma = float(m['gamma'])
    want = [t.clone() for t in dev.backup_select(B, V, gamma)]
    want_mx = dev.max_values(B, V)[0].clone()
    torch.cuda.synchronize()
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    outs = []
    for i in range(6):
        with torch.cuda.stream(s1 if i % 2 == 0 else s2):
            outs.append(dev.backup_select(B, V, gamma) if i % 3 else (dev.max_values(B, V)[0],))
    torch.cuda.synchronize()
    for o in outs:
        if len(o) == 1:
            assert torch.equal(o[0], want_mx)
        else:
            assert all(torch.equal(a, b) for a, b in zip(o, want) if a is not None)


@pytest.mark.parametrize('tag', ['tiger', 'grid4x4', 'hallway', 'olfactory_wrap'])
def test_chain_kernel_equals_multi_launch_chain(torch_cuda, tag):
    """The persistent one-launch chain (belief_chain_kernel: belief in shared memory, a dozen block barriers per step) == the
    multi-launch chain (one projection / leaf-sum / finish launch per step), bit for bit: FSVI trajectories with resets and with an
    impossible observation (NaN rows), Perseus walks with the observations drawn on the device."""
    dev, m, reach, _ = device_model(tag)
    A, O = m['rto'].shape[1], m['rto'].shape[2]
    rng = np.random.default_rng(8)
    n = 57
    b0 = m['start']
    acts = rng.integers(0, A, n).astype(np.int32)
    us = rng.random(n)
    # observations that are possible along the way (from a device walk), plus resets
    dev.set_option('chain_kernel', 0)
    walk0, obs0 = dev.perseus_walk(b0, acts, us, want_observations=True)
    resets = (rng.random(n) < 0.1)
    traj0 = dev.belief_trajectory(torch_cuda.as_tensor(b0).cuda(), acts, obs0.cpu().numpy(), resets)
    launches_multi = dev._lib.pbvi_last_launches(dev._h)
    dev.set_option('chain_kernel', 1)                 # one persistent block
    walk1, obs1 = dev.perseus_walk(b0, acts, us, want_observations=True)
    assert dev._lib.pbvi_last_launches(dev._h) == 1
    traj1 = dev.belief_trajectory(torch_cuda.as_tensor(b0).cuda(), acts, obs0.cpu().numpy(), resets)
    assert dev._lib.pbvi_last_launches(dev._h) == 1 and launches_multi == 3 * n
    assert torch_cuda.equal(obs0, obs1)
    assert np.array_equal(walk0.cpu().numpy(), walk1.cpu().numpy(), equal_nan=True)
    assert np.array_equal(traj0.cpu().numpy(), traj1.cpu().numpy(), equal_nan=True)
    # the cluster form (8 blocks, four cluster barriers per step)
    dev.set_option('chain_kernel', 2)
    walk2, obs2 = dev.perseus_walk(b0, acts, us, want_observations=True)
    traj2 = dev.belief_trajectory(torch_cuda.as_tensor(b0).cuda(), acts, obs0.cpu().numpy(), resets)
    assert torch_cuda.equal(obs0, obs2)
    assert np.array_equal(walk0.cpu().numpy(), walk2.cpu().numpy(), equal_nan=True)
    assert np.array_equal(traj0.cpu().numpy(), traj2.cpu().numpy(), equal_nan=True)
    # an impossible observation somewhere in the chain: NaN from there on, in both forms
    bad_obs = obs0.cpu().numpy().copy()
    out = orc.all_successors(reach, m['rto'], np.asarray(b0)[None, :])[0]
    impossible = [(a, o) for a in range(A) for o in range(O) if np.isnan(out[a, o]).all()]
    if impossible:
        acts2 = acts.copy()
        acts2[0], bad_obs[0] = impossible[0]
        dev.set_option('chain_kernel', 0)
        t0 = dev.belief_trajectory(torch_cuda.as_tensor(b0).cuda(), acts2, bad_obs, None)
        dev.set_option('chain_kernel', 2)
        t1 = dev.belief_trajectory(torch_cuda.as_tensor(b0).cuda(), acts2, bad_obs, None)
        assert np.isnan(t1.cpu().numpy()[0]).all()
        assert np.array_equal(t0.cpu().numpy(), t1.cpu().numpy(), equal_nan=True)
