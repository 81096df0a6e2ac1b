"""
Generates the golden fixtures in this directory by executing the UNMODIFIED reference
(/root/reference, via oracle/ref_loader.py) in the authoring container.  Run as

    python tests/golden/make_golden.py            # all fixtures
    python tests/golden/make_golden.py tiger      # only fixtures whose name contains "tiger"

The reference has no tests or golden vectors for backup / update / expand, so these
files ARE the pinning of oracle/pbvi_oracle.py: every array under a `ref_` key is an output of a
reference function, every other array is an input fed to it.  The GPU box has no /root/reference;
the tests there read only these .npz files.

Seeds: every stochastic reference call is preceded by np.random.seed(seed); random.seed(seed).
"""
import os
import random
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle.ref_loader import load_reference, quiet, REFERENCE_ROOT  # noqa: E402

ref = load_reference()
EXAMPLES = os.path.join(REFERENCE_ROOT, 'Experiments', 'Example Models')
OLF_DATA = os.path.join(REFERENCE_ROOT, 'Experiments', 'Olfactory Navigation', 'Data')
OLF_VF = os.path.join(REFERENCE_ROOT, 'Experiments', 'Olfactory Navigation', 'ValueFunctions', '20231113_182429_value_function.csv')


def seed_all(seed):
    np.random.seed(seed)
    random.seed(seed)


def save(name, **arrays):
    path = os.path.join(HERE, name + '.npz')
    np.savez_compressed(path, **arrays)
    print(f'  wrote {name}.npz ({os.path.getsize(path) / 1e6:.2f} MB)')


def model_arrays(m, gamma):
    d = dict(reach=m.reachable_states, probs=m.reachable_probabilities, obs_table=m.observation_table,
             rto=m.reachable_transitional_observation_table, rbar=m.expected_rewards_table,
             start=m.start_probabilities, gamma=np.float64(gamma), end_states=np.array(m.end_states, dtype=np.int64),
             min_reward=np.float64(m._min_reward), max_reward=np.float64(m._max_reward))
    if m.transition_table is not None and m.state_count <= 64:
        d['transition_table'] = m.transition_table
    if m.immediate_reward_table is not None and m.state_count <= 64:
        d['reward_table'] = m.immediate_reward_table
    return d


# --------------------------------------------------------------------------------------------
def olfactory_reference_model(wrap=True):
    """Notebook recipe, cells [3]-[15] of Olfactory_Alternation_Paper_Wrap.ipynb, executed against the reference Model."""
    import cv2
    import pandas as pd
    ppu = 30
    W, H = 12 * ppu + 1, 2 * ppu + 1
    S = W * H
    nose = pd.read_csv(os.path.join(OLF_DATA, 'statistics_abs_nose_3e6.dat'), sep=' ', skiprows=[0], names=list(range(320)), index_col=False).to_numpy().T
    ground = pd.read_csv(os.path.join(OLF_DATA, 'statistics_abs_ground_3e6.dat'), sep=' ', skiprows=[0], names=list(range(320)), index_col=False).to_numpy().T
    nose = cv2.resize(nose, dsize=((4 * ppu) + 1, ppu + 1))
    ground = cv2.resize(ground, dsize=((4 * ppu) + 1, ppu + 1))
    nose_p = np.zeros((H, W)); nose_p[15:46, 60:181] = nose
    ground_p = np.zeros((H, W)); ground_p[15:46, 60:181] = ground
    goal = 30 * W + 60
    obs = np.empty((S, 6, 3))
    obs[:, :5, 0] = (1 - ground_p.ravel()[:, None]); obs[:, :5, 1] = ground_p.ravel()[:, None]
    obs[:, 5, 0] = (1 - nose_p.ravel()); obs[:, 5, 1] = nose_p.ravel()
    obs[:, :, 2] = 0.0; obs[goal, :, :] = 0.0; obs[goal, :, 2] = 1.0
    grid = [[f's_{i}_{j}' for j in range(W)] for i in range(H)]
    reach = np.zeros((S, 6, 1), dtype=int)
    for s in range(S):
        if wrap:
            reach[s, 0, 0] = s - W if s - W >= 0 else (S - W) + s
            reach[s, 1, 0] = s + 1 if (s + 1) % W > 0 else s - (W - 1)
            reach[s, 2, 0] = s + W if s + W < S else s % W
            reach[s, 3, 0] = s - 1 if (s - 1) % W < (W - 1) else s + W - 1
        else:
            reach[s, 0, 0] = s - W if s - W >= 0 else s
            reach[s, 1, 0] = s + 1 if (s + 1) % W > 0 else s
            reach[s, 2, 0] = s + W if s + W < S else s
            reach[s, 3, 0] = s - 1 if (s - 1) % W < (W - 1) else s
        reach[s, 4, 0] = s
        reach[s, 5, 0] = s

    def reward_func(s, a, sn, o):
        return np.where(sn == goal, 1.0, 0.0)
    start = np.zeros((H, W)); start[15:46, 60:316] = 1.0; start /= np.sum(start)
    with quiet():
        m = ref.Model(states=grid, actions=['N', 'E', 'S', 'W', 'O_Ground', 'O_Air'], observations=['nothing', 'something', 'goal'],
                      reachable_states=reach, rewards=reward_func, observation_table=obs, end_states=[goal],
                      start_probabilities=start.ravel())
    return m, ground, nose


def per_belief_backup(solver, m, B, vf):
    """Reference backup called one belief at a time: exposes a*[b] and alpha_b without the set dedup."""
    acts, rows = [], []
    for b in B:
        with quiet():
            out = solver.backup(m, ref.BeliefSet(m, b[None, :].copy()), vf, append=False, belief_dominance_prune=False)
        assert len(out) == 1
        rows.append(out.alpha_vector_array[0]); acts.append(int(out.actions[0]))
    return np.array(rows), np.array(acts, dtype=np.int64)


def backup_fixture(name, m, gamma, B, vf, extra=None):
    solver = ref.PBVI_Solver(gamma=gamma, eps=1e-6, expand_function='fsvi')
    bs = ref.BeliefSet(m, B.copy())
    with quiet():
        full = solver.backup(m, bs, vf, append=False, belief_dominance_prune=False)
        full_arr, full_act = full.alpha_vector_array.copy(), np.array(full.actions).copy()
        app = solver.backup(m, bs, vf, append=True, belief_dominance_prune=False)
        app_arr, app_act = app.alpha_vector_array.copy(), np.array(app.actions).copy()
        dom = solver.backup(m, bs, vf, append=False, belief_dominance_prune=True)
        dom_arr = dom.alpha_vector_array.copy() if len(dom) else np.zeros((0, m.state_count))
        dom_act = np.array(dom.actions).copy() if len(dom) else np.zeros((0,), dtype=np.int64)
    rows, acts = per_belief_backup(solver, m, B, vf)
    d = dict(beliefs=B, alphas=vf.alpha_vector_array, alpha_actions=np.array(vf.actions),
             ref_row_alpha=rows, ref_row_action=acts,
             ref_vf_alpha=full_arr, ref_vf_action=full_act,
             ref_vf_append_alpha=app_arr, ref_vf_append_action=app_act,
             ref_vf_dom_alpha=dom_arr, ref_vf_dom_action=dom_act)
    if extra:
        d.update(extra)
    save(name, **d)


def solve_snapshots(m, gamma, flavour, expansions, growth, seed, **kw):
    seed_all(seed)
    solver = ref.PBVI_Solver(gamma=gamma, eps=1e-6, expand_function=flavour, **kw)
    with quiet():
        vf, hist = solver.solve(m, expansions=expansions, max_belief_growth=growth, history_tracking_level=2, print_progress=False)
    return solver, vf, hist


# --------------------------------------------------------------------------------------------
def gen_small_model(fname, tag, flavour='fsvi', expansions=6, growth=8, seed=0):
    print(f'[{tag}]')
    with quiet():
        m, s0 = ref.load_POMDP_file(os.path.join(EXAMPLES, fname))
    gamma = s0.gamma
    save(f'model_{tag}', **model_arrays(m, gamma))
    solver, vf, hist = solve_snapshots(m, gamma, flavour, expansions, growth, seed)
    # inputs: all beliefs explored, value function one step before the end (so the backup is non trivial)
    B = hist.belief_sets[-1].belief_array.copy()
    vprev = hist.value_functions[-2]
    backup_fixture(f'backup_{tag}', m, gamma, B, vprev)

    # belief updates for every (a,o) of the first beliefs  (Belief.update, src/pomdp.py:382-421)
    nb = min(6, B.shape[0])
    upd = np.empty((nb, m.action_count, m.observation_count, m.state_count))
    with np.errstate(all='ignore'):
        for i in range(nb):
            bel = ref.Belief(m, B[i].copy())
            for a in range(m.action_count):
                for o in range(m.observation_count):
                    upd[i, a, o] = bel.update(a, o).values
    # compute_change between the two last value functions on the explored beliefs (src/pomdp.py:2141-2169)
    chg = solver.compute_change(hist.value_functions[-1], vprev, hist.belief_sets[-1])
    # MDP value iteration (src/mdp.py:1442-1525)
    with quiet():
        vi_vf, vi_hist = ref.VI_Solver(gamma=gamma, eps=1e-6).solve(m, print_progress=False)
    save(f'misc_{tag}', beliefs=B[:nb], ref_updates=upd, change_beliefs=B,
         change_alphas_a=hist.value_functions[-1].alpha_vector_array, change_alphas_b=vprev.alpha_vector_array,
         ref_change=np.float64(chg),
         ref_vi_alpha=vi_vf.alpha_vector_array, ref_vi_action=np.array(vi_vf.actions), ref_vi_iters=np.int64(len(vi_hist.iteration_times)))
    return m, gamma


def gen_tiger_extras():
    """Expansion flavours, sawtooth, whole solves: only tiger-class models run every flavour in the reference."""
    print('[tiger extras]')
    with quiet():
        m, s0 = ref.load_POMDP_file(os.path.join(EXAMPLES, 'tiger.95.POMDP'))
    gamma = 0.95
    with quiet():
        mdp_vf, _ = ref.VI_Solver(gamma=gamma, eps=1e-6).solve(m, print_progress=False)
    out = dict(mdp_alpha=mdp_vf.alpha_vector_array, mdp_action=np.array(mdp_vf.actions))
    # a fixed belief set / value function to expand from
    solver, vf, hist = solve_snapshots(m, gamma, 'ssra', 4, 6, 1)
    B = hist.belief_sets[-1].belief_array.copy()
    V = hist.value_functions[-1]
    out.update(beliefs=B, alphas=V.alpha_vector_array, alpha_actions=np.array(V.actions))
    for flavour, kw in [('ra', {}), ('ssra', {}), ('ssga', {}), ('ssea', {}), ('ger', {}), ('fsvi', {'mdp_policy': mdp_vf}),
                        ('fsvi_eg', {'mdp_policy': mdp_vf}), ('perseus', {}), ('hsvi', {'mdp_policy': mdp_vf})]:
        sv = ref.PBVI_Solver(gamma=gamma, eps=1e-6, expand_function=flavour, **kw)
        seed_all(7)
        with quiet():
            nb = sv.expand(m, ref.BeliefSet(m, B.copy()), max_generation=5, value_function=V, **sv.expand_function_params)
        out[f'ref_expand_{flavour}'] = nb.belief_array.copy()
    # sawtooth (src/pomdp.py:873-895) after HSVI populated the upper bound
    ub = sv._upper_bound
    ub.update()
    q = np.array([[0.3, 0.7], [0.5, 0.5], [0.05, 0.95], [0.999, 0.001]])
    out.update(ub_beliefs=ub.belief_array.copy(), ub_values=ub.value_array.copy(), ub_corner=ub.corner_values.copy(), ub_queries=q,
               ref_ub_eval=np.array([ub.evaluate(ref.Belief(m, qq)) for qq in q]))
    # whole solves, one per flavour
    for flavour in ['ra', 'ssra', 'ssga', 'ssea', 'ger', 'fsvi', 'fsvi_eg', 'perseus', 'hsvi']:
        solver, vf, hist = solve_snapshots(m, gamma, flavour, 6, 8, 3)
        out[f'ref_solve_{flavour}_alpha'] = vf.alpha_vector_array.copy()
        out[f'ref_solve_{flavour}_action'] = np.array(vf.actions).copy()
        out[f'ref_solve_{flavour}_beliefs'] = hist.belief_sets[-1].belief_array.copy()
        out[f'ref_solve_{flavour}_vcounts'] = np.array(hist.alpha_vector_counts)
        out[f'ref_solve_{flavour}_bcounts'] = np.array(hist.beliefs_counts)
    save('tiger_extras', **out)


GRID_FLAVOURS = ['ra', 'ssra', 'ssga', 'fsvi', 'fsvi_eg', 'perseus', 'hsvi']     # SSEA / GER crash on grids in the reference (NaN successors)


def gen_grid_extras():
    """Expansion flavours and whole solves on the 4x4 no-loop grid (R = 1, end state, impossible (a, o) pairs): everything the
    reference can run there -- its SSEA / GER assert on the NaN successors of impossible observations."""
    print('[grid extras]')
    with quiet():
        m, s0 = ref.load_POMDP_file(os.path.join(EXAMPLES, '4x4.95-no_loop.POMDP'))
    gamma = s0.gamma
    with quiet():
        mdp_vf, _ = ref.VI_Solver(gamma=gamma, eps=1e-6).solve(m, print_progress=False)
    out = dict(mdp_alpha=mdp_vf.alpha_vector_array, mdp_action=np.array(mdp_vf.actions), gamma=np.float64(gamma))
    solver, vf, hist = solve_snapshots(m, gamma, 'perseus', 4, 6, 1)
    B = hist.belief_sets[-1].belief_array.copy()
    V = hist.value_functions[-1]
    out.update(beliefs=B, alphas=V.alpha_vector_array, alpha_actions=np.array(V.actions))
    for flavour in GRID_FLAVOURS:
        kw = {'mdp_policy': mdp_vf} if flavour in ('fsvi', 'fsvi_eg', 'hsvi') else {}
        sv = ref.PBVI_Solver(gamma=gamma, eps=1e-6, expand_function=flavour, **kw)
        seed_all(7)
        with quiet(), np.errstate(all='ignore'):
            nb = sv.expand(m, ref.BeliefSet(m, B.copy()), max_generation=5, value_function=V, **sv.expand_function_params)
        out[f'ref_expand_{flavour}'] = nb.belief_array.copy()
    for flavour in GRID_FLAVOURS:
        with np.errstate(all='ignore'):
            solver, vf, hist = solve_snapshots(m, gamma, flavour, 5, 8, 3)
        out[f'ref_solve_{flavour}_alpha'] = vf.alpha_vector_array.copy()
        out[f'ref_solve_{flavour}_action'] = np.array(vf.actions).copy()
        out[f'ref_solve_{flavour}_beliefs'] = hist.belief_sets[-1].belief_array.copy()
        out[f'ref_solve_{flavour}_vcounts'] = np.array(hist.alpha_vector_counts)
        out[f'ref_solve_{flavour}_bcounts'] = np.array(hist.beliefs_counts)
    save('grid_extras', **out)


def gen_olfactory():
    print('[olfactory]')
    m, ground, nose = olfactory_reference_model(wrap=True)
    save('olfactory_maps', ground=ground, nose=nose)
    gamma = 0.99
    # model tensors are large (RTO 3.2 MB) but compress well; store the checksum-relevant ones
    save('model_olfactory_wrap', reach=m.reachable_states.astype(np.int32), rto=m.reachable_transitional_observation_table,
         rbar=m.expected_rewards_table, start=m.start_probabilities, gamma=np.float64(gamma),
         end_states=np.array(m.end_states, dtype=np.int64))
    # FSVI run (as published shape, shortened): realistic trajectory beliefs + a grown value function
    with quiet():
        mdp_vf, mdp_hist = ref.VI_Solver(gamma=gamma, eps=1e-6).solve(m, print_progress=False)
    print('  VI iterations (wrap, eps 1e-6):', len(mdp_hist.iteration_times))
    solver, vf, hist = solve_snapshots(m, gamma, 'fsvi', 5, 40, 0, mdp_policy=mdp_vf)
    Ball = hist.belief_sets[-1].belief_array
    vprev = hist.value_functions[-2]
    print('  beliefs', Ball.shape, 'V', len(vprev))
    sel = np.linspace(0, Ball.shape[0] - 1, 20).astype(int)
    B = Ball[sel].copy()
    backup_fixture('backup_olfactory_wrap', m, gamma, B, vprev,
                   extra=dict(ref_vi_iters=np.int64(len(mdp_hist.iteration_times)),
                              ref_vi_vopt=np.max(mdp_vf.alpha_vector_array, axis=0)))
    # belief updates along the trajectory
    nb = 4
    pairs = [(0, 0), (1, 1), (4, 0), (5, 1), (2, 0), (3, 2)]
    upd = np.empty((nb, len(pairs), m.state_count))
    with np.errstate(all='ignore'):
        for i in range(nb):
            bel = ref.Belief(m, B[i * 3].copy())
            for j, (a, o) in enumerate(pairs):
                upd[i, j] = bel.update(a, o).values
    chg = solver.compute_change(hist.value_functions[-1], vprev, ref.BeliefSet(m, B.copy()))
    save('misc_olfactory_wrap', beliefs=B[0:nb * 3:3], pairs=np.array(pairs), ref_updates=upd,
         change_alphas_a=hist.value_functions[-1].alpha_vector_array[:30], change_alphas_b=vprev.alpha_vector_array,
         ref_change=np.float64(solver.compute_change(ref.ValueFunction(m, hist.value_functions[-1].alpha_vector_array[:30], np.array(hist.value_functions[-1].actions)[:30]), vprev, ref.BeliefSet(m, B.copy()))),
         change_beliefs=B)

    # known-answer test: the checked-in MDP solution of the NON-wrap model (VI_Solver(0.99, 1e-4), 460 iterations)
    m2, _, _ = olfactory_reference_model(wrap=False)
    import pandas as pd
    kat = pd.read_csv(OLF_VF).to_numpy()
    with quiet():
        vi2, h2 = ref.VI_Solver(gamma=0.99, eps=1e-4).solve(m2, print_progress=False)
    print('  non-wrap VI iterations', len(h2.iteration_times), 'max |ours(ref code) - csv|', np.max(np.abs(vi2.alpha_vector_array - kat[:, 1:])))
    save('olf_nowrap_vi_kat', kat_action=kat[:, 0].astype(np.int64), kat_alpha=kat[:, 1:], ref_vi_iters=np.int64(len(h2.iteration_times)))


def gen_synthetic():
    """BASELINE configs[4] recipe at the small end: R>1, many-to-one reach, through the reference Model ctor."""
    print('[synthetic]')
    S, A, O, R = 300, 4, 3, 3
    rng = np.random.default_rng(5)
    RS = rng.integers(0, S, (S, A, R))
    OT = rng.random((S, A, O)); OT /= OT.sum(2, keepdims=True)
    with quiet():
        m = ref.Model(states=S, actions=A, observations=O, reachable_states=RS, observation_table=OT, end_states=[S // 2])
    gamma = 0.95
    save('model_synth300', **model_arrays(m, gamma))
    B = np.zeros((24, S))
    for i in range(24):
        k = [S, 40, 5, 1][i % 4]
        idx = rng.choice(S, k, replace=False)
        B[i, idx] = rng.dirichlet(np.ones(k))
    B[3] = B[2]            # duplicate belief -> duplicate alpha rows
    alph = rng.random((17, S))
    alph[5] = alph[1]      # duplicated alpha (ValueFunction dedups it)
    acts = rng.integers(0, A, 17)
    with quiet():
        vf = ref.ValueFunction(m, alph, acts)
    backup_fixture('backup_synth300', m, gamma, B, vf)


def gen_simulations():
    """Agent roll-outs (src/pomdp.py:2953-3380): seeded single simulations and the vectorised n-parallel simulation on an
    R = 1 model (the reference's vectorised next-state gather is only well defined there), plus test_n_simulations."""
    print('[simulations]')
    out = {}
    for fname, tag in [('4x4.95-no_loop.POMDP', 'grid4x4_noloop'), ('tiger.95.POMDP', 'tiger')]:
        with quiet():
            m, s0 = ref.load_POMDP_file(os.path.join(EXAMPLES, fname))
        g = dict(np.load(os.path.join(HERE, f'backup_{tag}.npz')))
        with quiet():
            vf = ref.ValueFunction(m, g['alphas'], g['alpha_actions'])
        agent = ref.Agent(m, vf)
        seed_all(11)
        with quiet():
            h = agent.simulate(max_steps=25, print_progress=False, print_stats=False)
        out[f'{tag}_sim_states'] = np.array(h.states)
        out[f'{tag}_sim_actions'] = np.array(h.actions)
        out[f'{tag}_sim_observations'] = np.array(h.observations)
        out[f'{tag}_sim_rewards'] = np.array(h.rewards, dtype=float)
        out[f'{tag}_sim_last_belief'] = h.beliefs[-1].values
        seed_all(12)
        with quiet():
            rs, hs = agent.run_n_simulations(n=4, max_steps=15, print_progress=False, print_stats=False)
        out[f'{tag}_nsim_totals'] = np.array(rs, dtype=float)
        out[f'{tag}_nsim_lengths'] = np.array([len(x) for x in hs])
    # vectorised simulation: needs a reward FUNCTION (SimulationSet.run_actions calls model.immediate_reward_function, :2935) and R = 1
    H, W = 5, 7
    S = H * W
    s_idx = np.arange(S)
    row, col = s_idx // W, s_idx % W
    reach = np.zeros((S, 5, 1), dtype=int)
    reach[:, 0, 0] = ((row - 1) % H) * W + col
    reach[:, 1, 0] = row * W + (col + 1) % W
    reach[:, 2, 0] = ((row + 1) % H) * W + col
    reach[:, 3, 0] = row * W + (col - 1) % W
    reach[:, 4, 0] = s_idx
    goal = 2 * W + 1
    rng = np.random.default_rng(3)
    p_detect = rng.random(S) * 0.6
    obs = np.zeros((S, 5, 3))
    obs[:, :, 0] = 1 - p_detect[:, None]
    obs[:, :, 1] = p_detect[:, None]
    obs[goal, :, :] = [0.0, 0.0, 1.0]
    start = np.ones(S); start[goal] = 0; start /= start.sum()

    def reward_func(s, a, sn, o):
        return np.where(sn == goal, 1.0, 0.0)
    with quiet():
        m = ref.Model(states=[[f's_{i}_{j}' for j in range(W)] for i in range(H)], actions=5, observations=3, reachable_states=reach,
                      rewards=reward_func, observation_table=obs, end_states=[goal], start_probabilities=start)
    seed_all(5)
    solver = ref.FSVI_Solver(gamma=0.95, eps=1e-6)
    with quiet():
        vf, _ = solver.solve(m, expansions=12, max_belief_growth=15, print_progress=False)
    out.update(grid_reach=reach, grid_obs=obs, grid_start=start, grid_goal=np.int64(goal), grid_shape=np.array([H, W]),
               grid_alphas=vf.alpha_vector_array, grid_alpha_actions=np.array(vf.actions))
    agent = ref.Agent(m, vf)
    seed_all(13)
    with quiet():
        rs, hs = agent.run_n_simulations_parallel(n=24, max_steps=30, print_progress=False, print_stats=False)
    out['grid_par_totals'] = np.array(rs, dtype=float)
    out['grid_par_lengths'] = np.array([len(x) for x in hs])
    out['grid_par_states0'] = np.array(hs[0].states)
    out['grid_par_actions_last'] = np.array(hs[-1].actions)
    out['grid_par_observations_last'] = np.array(hs[-1].observations)
    seed_all(14)
    with quiet():
        starts, done_at, rew, drew = solver.test_n_simulations(m, vf, n=16, horizon=20)
    out['grid_test_starts'] = np.array(starts)
    out['grid_test_done_at'] = np.array(done_at)
    out['grid_test_rewards'] = np.array(rew, dtype=float)
    out['grid_test_discounted'] = np.array(drew, dtype=float)
    seed_all(15)
    with quiet():
        h = agent.simulate(max_steps=25, print_progress=False, print_stats=False)
    out['grid_sim_states'] = np.array(h.states)
    out['grid_sim_actions'] = np.array(h.actions)
    out['grid_sim_observations'] = np.array(h.observations)
    save('simulations', **out)


def gen_io():
    """
    On-disk formats (src/mdp.py:909-1036): files WRITTEN BY THE UNMODIFIED REFERENCE -- `ValueFunction.save` (csv, csv.gzip) and
    `save_parquet` of a small value function (the reference's own backup of the 4x4 no_loop fixture) -- plus the one value-function
    artefact the reference ships, gzip-compressed as its own `save(compress=True)` would (the loader keys on '.gzip' in the name).
    """
    import gzip
    import shutil
    print('[io]')
    out_dir = os.path.join(HERE, 'io')
    os.makedirs(out_dir, exist_ok=True)
    with quiet():
        m, s0 = ref.load_POMDP_file(os.path.join(EXAMPLES, '4x4.95-no_loop.POMDP'))
    g = dict(np.load(os.path.join(HERE, 'backup_grid4x4_noloop.npz')))
    vf = ref.ValueFunction(m, g['alphas'], g['alpha_actions'])
    with quiet():
        vf.save(path=out_dir, file_name='ref_saved_grid4x4_noloop.csv')
        vf.save(path=out_dir, file_name='ref_saved_grid4x4_noloop_gz.csv', compress=True)
        vf.save_parquet(path=out_dir, file_name='ref_saved_grid4x4_noloop.parquet')
    # what the reference's own loaders return for its own files (pandas' default csv float parser is not round-trip exact: the
    # reference's csv round trip moves some entries by one ulp; parquet is exact)
    back_csv = ref.ValueFunction.load_from_file(os.path.join(out_dir, 'ref_saved_grid4x4_noloop.csv'), m)
    back_gz = ref.ValueFunction.load_from_file(os.path.join(out_dir, 'ref_saved_grid4x4_noloop_gz.csv.gzip'), m)
    back_pq = ref.ValueFunction.load_from_parquet(os.path.join(out_dir, 'ref_saved_grid4x4_noloop.parquet'), m)
    assert np.array_equal(back_pq.alpha_vector_array, vf.alpha_vector_array)
    assert np.array_equal(back_csv.alpha_vector_array, back_gz.alpha_vector_array)
    print('  reference csv round trip: max |loaded - saved| =', np.max(np.abs(np.asarray(back_csv.alpha_vector_array) - np.asarray(vf.alpha_vector_array))))
    save('io_grid4x4_noloop', alphas=np.asarray(vf.alpha_vector_array), actions=np.asarray(vf.actions, dtype=np.int64),
         state_labels=np.array(m.state_labels), ref_loaded_csv_alphas=np.asarray(back_csv.alpha_vector_array),
         ref_loaded_csv_actions=np.asarray(back_csv.actions, dtype=np.int64), ref_loaded_parquet_alphas=np.asarray(back_pq.alpha_vector_array),
         ref_loaded_parquet_actions=np.asarray(back_pq.actions, dtype=np.int64))
    with open(OLF_VF, 'rb') as f_in, gzip.GzipFile(os.path.join(out_dir, 'ref_20231113_182429_value_function.csv.gzip'), 'wb', mtime=0) as f_out:
        shutil.copyfileobj(f_in, f_out)
    for f in sorted(os.listdir(out_dir)):
        print(f'  wrote io/{f} ({os.path.getsize(os.path.join(out_dir, f)) / 1e3:.1f} kB)')


if __name__ == '__main__':
    want = sys.argv[1:] or ['']
    def on(tag):
        return any(w in tag for w in want)
    if on('tiger'):
        gen_small_model('tiger.95.POMDP', 'tiger', flavour='ssra', expansions=5, growth=8)
        gen_tiger_extras()
    if on('4x4'):
        gen_small_model('4x4.95.POMDP', 'grid4x4', flavour='fsvi', expansions=6, growth=8)
        gen_small_model('4x4.95-no_loop.POMDP', 'grid4x4_noloop', flavour='perseus', expansions=6, growth=8)
    if on('grid_extras'):
        gen_grid_extras()
    # the other example models the reference's reader accepts (its parser rejects parr95, saci-s12-a6-z5 and shuttle; cit is 284 states)
    if on('cheese'):
        gen_small_model('cheese.95.POMDP', 'cheese', flavour='fsvi', expansions=6, growth=8)
    if on('network'):
        gen_small_model('network.95.POMDP', 'network', flavour='perseus', expansions=6, growth=8)
    if on('hanks'):
        gen_small_model('hanks.95.POMDP', 'hanks', flavour='ssra', expansions=6, growth=8)
    if on('grid4x3'):
        gen_small_model('4x3.95.POMDP', 'grid4x3', flavour='fsvi', expansions=6, growth=8)
    if on('maze4x5x2'):
        gen_small_model('4x5x2.95.POMDP', 'maze4x5x2', flavour='perseus', expansions=6, growth=8)
    if on('cit'):
        gen_small_model('cit.POMDP', 'cit', flavour='fsvi', expansions=4, growth=10)
    if on('tigergrid'):
        gen_small_model('tiger-grid.POMDP', 'tigergrid', flavour='fsvi', expansions=6, growth=10)
    if on('hallway'):
        gen_small_model('hallway.POMDP', 'hallway', flavour='fsvi', expansions=5, growth=10)
    if on('synth'):
        gen_synthetic()
    if on('olfactory'):
        gen_olfactory()
    if on('simulations'):
        gen_simulations()
    if on('io'):
        gen_io()
