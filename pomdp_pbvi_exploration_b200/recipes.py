"""
Model recipes used as benchmark / parity inputs.

These are *inputs* to the engine, rebuilt from the notebooks' recipes -- the reference's pickled
models are missing blobs (.MISSING_LARGE_BLOBS) -- not part of the hot path:

* `tiger_model`            Cassandra's tiger.95 (S2 A3 O2 R2), the reference's CPU-runnable config.
* `olfactory_wrap_model`   the 61x361 = 22021-state toroidal olfactory-navigation POMDP of
                           `Experiments/Olfactory Navigation/Olfactory_Alternation_Paper_Wrap.ipynb[3-15]`
                           (6 actions, 3 observations, deterministic moves => R=1).
* `synthetic_sparse_model` the random sparse POMDP recipe of BASELINE.json configs[4].
* `sea_robin_model`        the 63 555-state sea-robin tank POMDP (A = 16, O = 2), the reference's largest workload.
"""
from __future__ import annotations

import os

import numpy as np

from .model import Model

_GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')


def tiger_model() -> Model:
    """tiger.95: listen / open-left / open-right; hearing accuracy 0.85; +10 / -100 / -1."""
    T = np.empty((2, 3, 2))
    T[:, 0, :] = np.eye(2)
    T[:, 1:, :] = 0.5
    Obs = np.full((2, 3, 2), 0.5)
    Obs[0, 0] = [0.85, 0.15]
    Obs[1, 0] = [0.15, 0.85]
    rew_sa = np.array([[-1.0, -100.0, 10.0], [-1.0, 10.0, -100.0]])
    Rw = np.broadcast_to(rew_sa[:, :, None, None], (2, 3, 2, 2)).copy()
    return Model(states=['tiger-left', 'tiger-right'], actions=['listen', 'open-left', 'open-right'],
                 observations=['tiger-left', 'tiger-right'], transitions=T, rewards=Rw, observation_table=Obs,
                 start_probabilities=[0.5, 0.5])


def load_olfactory_maps(path: str | None = None):
    """
    The two plume detection-probability maps (ground, nose/air), each 31x121, as produced by the notebook's
    `cv2.resize(data.T, dsize=(121, 31))` of `Data/statistics_abs_{ground,nose}_3e6.dat` (cell [4]).
    Stored as a derived fixture (`tests/golden/olfactory_maps.npz`, written by tests/golden/make_golden.py)
    because neither the raw data nor the reference travel to the GPU box.
    """
    path = path or os.path.join(_GOLDEN_DIR, 'olfactory_maps.npz')
    z = np.load(path)
    return z['ground'], z['nose']


def olfactory_wrap_model(ground_map: np.ndarray | None = None, nose_map: np.ndarray | None = None,
                         points_per_unit: int = 30, wrap: bool = True) -> Model:
    """
    Olfactory navigation POMDP (notebook cells [3]-[15]).  Grid (2*ppu+1) x (12*ppu+1); actions N,E,S,W,
    sniff-ground, sniff-air; observations nothing / something / goal; the plume maps are placed at
    rows [ppu/2, ppu/2 + ppu], cols [2*ppu, 6*ppu]; source (goal) at (ppu, 2*ppu); reward 1 on landing on
    the goal; start belief uniform over rows [ppu/2, 1.5*ppu], cols [2*ppu, 10.5*ppu].
    `wrap=False` gives the non-toroidal variant of `Olfactory_Alternation_Paper.ipynb` (moves off the grid stay put).
    With ppu != 30 the maps are resampled by nearest neighbour (used for reduced-size test models only).
    """
    if ground_map is None or nose_map is None:
        ground_map, nose_map = load_olfactory_maps()
    ppu = points_per_unit
    H, W = 2 * ppu + 1, 12 * ppu + 1
    S = H * W
    r0, c0 = ppu // 2, 2 * ppu
    mh, mw = ppu + 1, 4 * ppu + 1
    if ground_map.shape != (mh, mw):
        ri = np.round(np.linspace(0, ground_map.shape[0] - 1, mh)).astype(int)
        ci = np.round(np.linspace(0, ground_map.shape[1] - 1, mw)).astype(int)
        ground_map = ground_map[np.ix_(ri, ci)]
        nose_map = nose_map[np.ix_(ri, ci)]
    ground = np.zeros((H, W)); ground[r0:r0 + mh, c0:c0 + mw] = ground_map
    nose = np.zeros((H, W)); nose[r0:r0 + mh, c0:c0 + mw] = nose_map
    goal = ppu * W + 2 * ppu

    obs = np.empty((S, 6, 3))
    obs[:, :5, 0] = (1 - ground.ravel()[:, None])
    obs[:, :5, 1] = ground.ravel()[:, None]
    obs[:, 5, 0] = (1 - nose.ravel())
    obs[:, 5, 1] = nose.ravel()
    obs[:, :, 2] = 0.0
    obs[goal, :, :] = 0.0
    obs[goal, :, 2] = 1.0

    s = np.arange(S)
    row, col = s // W, s % W
    reach = np.zeros((S, 6, 1), dtype=int)
    if wrap:
        reach[:, 0, 0] = ((row - 1) % H) * W + col
        reach[:, 1, 0] = row * W + (col + 1) % W
        reach[:, 2, 0] = ((row + 1) % H) * W + col
        reach[:, 3, 0] = row * W + (col - 1) % W
    else:
        reach[:, 0, 0] = np.where(row > 0, s - W, s)
        reach[:, 1, 0] = np.where(col < W - 1, s + 1, s)
        reach[:, 2, 0] = np.where(row < H - 1, s + W, s)
        reach[:, 3, 0] = np.where(col > 0, s - 1, s)
    reach[:, 4, 0] = s
    reach[:, 5, 0] = s

    start = np.zeros((H, W))
    start[r0:r0 + mh, c0:c0 + int(8.5 * ppu) + 1] = 1.0
    start /= np.sum(start)

    def reward_func(s_, a_, sn_, o_):
        return np.where(sn_ == goal, 1.0, 0.0)

    grid_labels = [[f's_{i}_{j}' for j in range(W)] for i in range(H)]
    return Model(states=grid_labels, actions=['N', 'E', 'S', 'W', 'O_Ground', 'O_Air'],
                 observations=['nothing', 'something', 'goal'], reachable_states=reach, rewards=reward_func,
                 observation_table=obs, end_states=[goal], start_probabilities=start.ravel())


def sea_robin_model(tank_size=(111, 142), swim_radius: float = 8.5, source_radius: float = 10.5) -> Model:
    """
    The sea-robin "walk + swim" tank POMDP of `Experiments/Sea Robins/Sea_Robins_Swim_Walk.ipynb[2-12]` -- the reference's largest
    workload (S = 223 x 285 = 63 555, A = 16, O = 2, R = 1; its FSVI solve runs out of memory at |V| = 1364 because
    Gamma = 8*A*O*V*S bytes = 22 GB).  State = agent position relative to the source on a toroidal grid of
    (2*tank + 1); 4 walking moves (1 cell) that sense ('something' inside the source disc, else 'nothing') and 12 swimming
    moves (the distinct lattice points at distance ~8.5: r in {0, 4}) that always observe 'nothing'; the end state is the centre.
    """
    tank = np.array(tank_size)
    H, W = (tank * 2 + 1).tolist()
    S = H * W
    walking = [np.array([1, 0]), np.array([0, 1]), np.array([-1, 0]), np.array([0, -1])]
    moves = []
    for r in [0, 4]:
        d = int(np.floor(np.sqrt(swim_radius ** 2 - r ** 2)))
        moves += [np.array([-d, r]), np.array([d, r]), np.array([-d, -r]), np.array([d, -r]),
                  np.array([r, -d]), np.array([r, d]), np.array([-r, -d]), np.array([-r, d])]
    swimming = list({m.tobytes(): m for m in moves}.values())          # dict insertion order, like the notebook
    all_moves = walking + swimming
    A = len(all_moves)
    xs, ys = np.meshgrid(np.arange(H), np.arange(W), indexing='ij')
    like = (((xs - tank[0]) ** 2 + (ys - tank[1]) ** 2) <= source_radius ** 2).astype(float).ravel()
    reach = np.empty((S, A, 1), dtype=int)
    for a, mv in enumerate(all_moves):
        nx, ny = (xs + mv[0]) % H, (ys + mv[1]) % W
        reach[:, a, 0] = (nx * W + ny).ravel()
    obs = np.empty((S, A, 2))
    obs[:, :4, 0], obs[:, :4, 1] = like[:, None], 1.0 - like[:, None]
    obs[:, 4:, 0], obs[:, 4:, 1] = 0.0, 1.0
    grid_labels = [[f's_{i}_{j}' for j in range(W)] for i in range(H)]
    return Model(states=grid_labels, actions=[f'm_{int(m[0])}_{int(m[1])}' for m in all_moves], observations=['something', 'nothing'],
                 reachable_states=reach, observation_table=obs, end_states=[int(W * (H - 1) / 2 + (W - 1) / 2)])


def synthetic_sparse_model(S: int, A: int, O: int, R: int, seed: int = 0) -> Model:
    """Random sparse POMDP (BASELINE.json configs[4]): uniform 1/R transitions to
    `rng.integers(0,S,(S,A,R))`, Dirichlet-free normalised random observation table, reward 1 on reaching S//2."""
    rng = np.random.default_rng(seed)
    reach = rng.integers(0, S, (S, A, R))
    obs = rng.random((S, A, O))
    obs /= obs.sum(2, keepdims=True)
    return Model(states=S, actions=A, observations=O, reachable_states=reach, observation_table=obs, end_states=[S // 2])


def perseus_walk_beliefs(model: Model, n: int, seed: int = 0, restart: int = 100) -> np.ndarray:
    """
    Synthetic belief points for benchmarks / tests: the random walk in belief space of the reference's `expand_perseus`
    (src/pomdp.py:2010-2056: a uniform, o ~ P(o|b,a), b <- update(b,a,o)), restarted from the start belief every
    `restart` steps.  Host NumPy input synthesis with its own RNG stream (`default_rng(seed)`); not part of the solver.
    """
    rng = np.random.default_rng(seed)
    S, A, O = model.state_count, model.action_count, model.observation_count
    reach = model.reachable_states
    rto = model.reachable_transitional_observation_table
    out = np.empty((n, S))
    b = model.start_probabilities
    for i in range(n):
        if i % restart == 0:
            b = model.start_probabilities
        a = int(rng.integers(A))
        w = rto[:, a, :, :] * b[:, None, None]                    # [S,O,R]
        po = w.sum(axis=(0, 2))
        o = int(rng.choice(O, p=po / po.sum()))
        nb = np.bincount(reach[:, a, :].ravel(), weights=w[:, o, :].ravel(), minlength=S)
        b = nb / nb.sum()
        out[i] = b
    return out
