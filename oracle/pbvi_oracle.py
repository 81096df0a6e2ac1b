"""
TEST INFRASTRUCTURE ONLY -- CPU oracle for the PBVI hot path (NumPy restatement of the reference).

This file restates, in plain NumPy on bare arrays, the arithmetic of the reference's hot path
(PimLb/POMDP_PBVI_Exploration, `src/pomdp.py` / `src/mdp.py`), function by function, each citing
the reference lines it follows.  It is the *checker* for the CUDA engine and the `cpu_baseline`
leg of `bench.py`.  It is never imported by the product package
(`pomdp_pbvi_exploration_b200/`); only `tests/`, `__graft_entry__.smoke()` and `bench.py` (for
`cpu_baseline` / `--impl reference`) may import it.

Parity pinning: the reference ships no tests and no golden vectors for backup / update / expand
(SURVEY.md section 4), so this oracle is pinned by running the *unmodified reference itself* in the
authoring container (`oracle/ref_loader.py`) on identical inputs:
  * `tests/golden/make_golden.py` dumps reference inputs/outputs into `tests/golden/*.npz`;
  * `tests/test_oracle_vs_golden.py` checks every function here against those fixtures (bit-exact
    where the reference's operation order is reproducible, see each docstring);
  * the one artefact the reference does pin -- the MDP value-iteration solution
    `Experiments/Olfactory Navigation/ValueFunctions/20231113_182429_value_function.csv` -- is a
    known-answer test for `vi_solve` (fixture `olf_nowrap_vi_kat.npz`).

Array conventions are the reference's (src/pomdp.py:112-122): `reach` = reachable_states [S,A,R] int,
`rto` = reachable_transitional_observation_table [S,A,O,R] f64, `rbar` = expected_rewards_table [S,A].
The arithmetic deliberately uses the same NumPy primitives as the reference (einsum / tensordot /
fancy indexing / bincount) so that (a) summation orders match and (b) timing it is a fair stand-in
for the reference's CPU path.
"""
from __future__ import annotations

import numpy as np

# --------------------------------------------------------------------------------------------
# Model tensors (row a5)
# --------------------------------------------------------------------------------------------

def reachable_probabilities_uniform(reach: np.ndarray) -> np.ndarray:
    """P[s,a,r] when neither a transition table nor function is given: 1/R everywhere (src/mdp.py:345-346)."""
    return np.full(reach.shape, 1.0 / reach.shape[2])


def reachable_probabilities_from_table(transition_table: np.ndarray, reach: np.ndarray) -> np.ndarray:
    """P[s,a,r] = T[s,a,reach[s,a,r]] (src/mdp.py:347-348)."""
    S, A, _ = reach.shape
    return transition_table[np.arange(S)[:, None, None], np.arange(A)[None, :, None], reach]


def derive_reachable_states(transition_table: np.ndarray) -> np.ndarray:
    """
    Reachable-state table from a dense T: argwhere(T[s,a,:] > 0), short rows padded with the
    smallest state ids not already present (src/mdp.py:306-335).
    """
    S, A, _ = transition_table.shape
    lists = [[np.argwhere(transition_table[s, a, :] > 0)[:, 0].tolist() for a in range(A)] for s in range(S)]
    R = max(len(l) for row in lists for l in row)
    for row in lists:
        for l in row:
            cand = 0
            while len(l) < R:
                if cand not in l:
                    l.append(cand)
                cand += 1
    return np.array(lists, dtype=int)


def build_rto(reach: np.ndarray, probs: np.ndarray, observation_table: np.ndarray) -> np.ndarray:
    """
    RTO[s,a,o,r] = P[s,a,r] * Obs[reach[s,a,r], a, o] -- the observation table is indexed by the
    LANDING state (src/pomdp.py:201-202).
    """
    S, A, R = reach.shape
    O = observation_table.shape[2]
    landing_obs = observation_table[reach[:, :, None, :], np.arange(A)[None, :, None, None], np.arange(O)[None, None, :, None]]
    return np.einsum('sar,saor->saor', probs, landing_obs)


def expected_rewards(rto: np.ndarray, reachable_rewards: np.ndarray) -> np.ndarray:
    """Rbar[s,a] = sum_{o,r} RTO[s,a,o,r] * rew[s,a,r,o] (src/pomdp.py:251)."""
    return np.einsum('saor,saro->sa', rto, reachable_rewards)


def end_state_reachable_rewards(reach: np.ndarray, n_obs: int, end_states, end_actions=()) -> np.ndarray:
    """rew[s,a,r,o] = 1 on landing in an end state or playing an end action (src/pomdp.py:239-246,257-258)."""
    S, A, R = reach.shape
    landing = np.isin(reach, list(end_states))
    act = np.isin(np.arange(A), list(end_actions))[None, :, None]
    base = (landing | act).astype(int)
    return np.repeat(base[:, :, :, None], n_obs, axis=3).astype(float)


# --------------------------------------------------------------------------------------------
# Backup (rows a1-a4)
# --------------------------------------------------------------------------------------------

def gamma_projection(reach: np.ndarray, rto: np.ndarray, alphas: np.ndarray, gamma: float) -> np.ndarray:
    """
    Gamma[a,o,v,s] = gamma * sum_r RTO[s,a,o,r] * alpha[v, reach[s,a,r]]  (src/pomdp.py:1485-1491).
    The multiplication by gamma happens AFTER the r-sum, as in the reference.
    """
    V = alphas.shape[0]
    alpha_at_reach = alphas[np.arange(V)[:, None, None, None], reach[None, :, :, :]]          # [V,S,A,R]
    return gamma * np.einsum('saor,vsar->aovs', rto, alpha_at_reach)


def backup(reach: np.ndarray, rto: np.ndarray, rbar: np.ndarray, gamma: float,
           beliefs: np.ndarray, alphas: np.ndarray, return_scores: bool = False) -> dict:
    """
    Point-based backup without dedup (src/pomdp.py:1485-1506).

    Returns dict with
      v_star [B,A,O]  first-index argmax over v of belief . Gamma[a,o,v]     (:1495)
      a_star [B]      first-index argmax over a of belief . alpha_a          (:1505)
      alpha  [B,S]    alpha_a[b, a_star[b], :]                               (:1506)
      values [B,A]    belief . alpha_a  (the quantities a_star is taken over)
      scores [B,A,O,V] (optional) belief . Gamma
    """
    S, A, R = reach.shape
    O = rto.shape[2]
    G = gamma_projection(reach, rto, alphas, gamma)                                           # [A,O,V,S]
    scores = np.tensordot(beliefs, G, (1, 3))                                                 # [B,A,O,V]
    v_star = np.argmax(scores, axis=3)
    best_per_o = G[np.arange(A)[None, :, None, None], np.arange(O)[None, None, :, None],
                   v_star[:, :, :, None], np.arange(S)[None, None, None, :]]                  # [B,A,O,S]
    alpha_a = rbar.T + np.sum(best_per_o, axis=2)                                             # [B,A,S]
    values = np.einsum('bas,bs->ba', alpha_a, beliefs)
    a_star = np.argmax(values, axis=1)
    alpha = np.take_along_axis(alpha_a, a_star[:, None, None], axis=1)[:, 0, :]
    out = dict(v_star=v_star, a_star=a_star, alpha=alpha, values=values)
    if return_scores:
        out['scores'] = scores
    return out


def backup_chunked(reach, rto, rbar, gamma, beliefs, alphas, chunk: int = 256) -> dict:
    """Same as `backup` but over belief chunks (rows are independent given V); bounds the Gamma* temporary."""
    parts = [backup(reach, rto, rbar, gamma, beliefs[i:i + chunk], alphas) for i in range(0, beliefs.shape[0], chunk)]
    return {k: np.concatenate([p[k] for p in parts], axis=0) for k in ('v_star', 'a_star', 'alpha', 'values')}


def belief_dominance_filter(beliefs: np.ndarray, new_alpha: np.ndarray, old_alphas: np.ndarray) -> np.ndarray:
    """keep[b] = belief.new_alpha_b > max_v belief.alpha_v, strict (src/pomdp.py:1509-1512)."""
    best_new = np.sum(beliefs * new_alpha, axis=1)
    best_old = np.max(np.matmul(beliefs, old_alphas.T), axis=1)
    return best_new > best_old


# --------------------------------------------------------------------------------------------
# Set semantics on raw bytes (row a7)
# --------------------------------------------------------------------------------------------

def dedup_rows(rows: np.ndarray, actions: np.ndarray):
    """
    ValueFunction constructor dedup (src/mdp.py:659-669): dict keyed by `values.tobytes()` built in row
    order => surviving POSITION is the first occurrence, surviving ACTION is the last occurrence's.
    Returns (rows_out, actions_out, first_index) with first_index into the input.
    """
    table = {}
    for i in range(rows.shape[0]):
        key = rows[i].tobytes()
        if key in table:
            table[key][1] = int(actions[i])
        else:
            table[key] = [i, int(actions[i])]
    first = np.array([v[0] for v in table.values()], dtype=np.int64)
    acts = np.array([v[1] for v in table.values()], dtype=np.int64)
    return rows[first] if len(first) else rows[:0], acts, first


def extend_union(new_rows, new_actions, old_rows, old_actions):
    """
    `new.extend(old)` (src/mdp.py:773-774): dict.update => new rows keep their order, old rows not already
    present are appended in old order, and on a byte collision the OLD AlphaVector (old action) replaces
    the value in place.  Inputs are assumed already deduped (both are ValueFunctions in the reference).
    """
    table = {}
    for i in range(new_rows.shape[0]):
        table[new_rows[i].tobytes()] = (new_rows[i], int(new_actions[i]))
    for i in range(old_rows.shape[0]):
        table[old_rows[i].tobytes()] = (old_rows[i], int(old_actions[i]))
    rows = np.array([v[0] for v in table.values()]).reshape(-1, new_rows.shape[1] if new_rows.ndim == 2 else old_rows.shape[1])
    acts = np.array([v[1] for v in table.values()], dtype=np.int64)
    return rows, acts


def belief_union(a_rows: np.ndarray, b_rows: np.ndarray) -> np.ndarray:
    """BeliefSet.union (src/pomdp.py:600-603): self's unique rows in order, then unseen rows of other."""
    table = {}
    for r in a_rows:
        table[r.tobytes()] = r
    for r in b_rows:
        table[r.tobytes()] = r          # same bytes: value replaced by an identical row, position kept
    return np.array(list(table.values())).reshape(-1, a_rows.shape[1])


# --------------------------------------------------------------------------------------------
# Belief update (row a6)
# --------------------------------------------------------------------------------------------

def belief_update(reach: np.ndarray, rto: np.ndarray, belief: np.ndarray, a: int, o: int, normalise: bool = True) -> np.ndarray:
    """
    b'[s'] = sum_{(s,r): reach[s,a,r]=s'} RTO[s,a,o,r] * b[s], then b' /= sum(b') (src/pomdp.py:405-411).
    bincount accumulates in flattened (s,r) order; an impossible observation gives 0/0 = NaN (no guard).
    """
    S = reach.shape[0]
    w = rto[:, a, o, :] * belief[:, None]
    nb = np.bincount(reach[:, a, :].flatten(), weights=w.flatten(), minlength=S)
    if normalise:
        with np.errstate(divide='ignore', invalid='ignore'):
            nb = nb / np.sum(nb)
    return nb


def belief_update_batch(reach, rto, beliefs, actions, observations, normalise: bool = True) -> np.ndarray:
    """Row-wise `belief_update` (the reference's batched twin is src/pomdp.py:1415-1419)."""
    return np.stack([belief_update(reach, rto, beliefs[i], int(actions[i]), int(observations[i]), normalise)
                     for i in range(beliefs.shape[0])]) if beliefs.shape[0] else beliefs.copy()


def observation_probabilities(rto: np.ndarray, belief: np.ndarray, a: int) -> np.ndarray:
    """P(o | b, a) = einsum('sor,s->o', RTO[:,a,:,:], b) (src/pomdp.py:1814, 2046)."""
    return np.einsum('sor,s->o', rto[:, a, :, :], belief)


def all_successors(reach, rto, beliefs) -> np.ndarray:
    """succ[b,a,o,:] = update(b,a,o) for every triple (src/pomdp.py:1679, 1732); NaN rows where P(o|b,a)=0."""
    B = beliefs.shape[0]
    S, A, R = reach.shape
    O = rto.shape[2]
    out = np.empty((B, A, O, S))
    for b in range(B):
        for a in range(A):
            for o in range(O):
                out[b, a, o] = belief_update(reach, rto, beliefs[b], a, o)
    return out


# --------------------------------------------------------------------------------------------
# NumPy's float64 pairwise summation, restated (what `np.sum` does on a contiguous 1-D array).
# The CUDA belief normalisation reproduces this tree so that b' is bit-identical to the reference's.
# --------------------------------------------------------------------------------------------

def numpy_pairwise_sum(a: np.ndarray) -> float:
    """numpy/_core/src/umath/loops_utils.h.src `@TYPE@_pairwise_sum` (NumPy 1.26 / 2.x), unit stride."""
    n = a.shape[0]
    if n < 8:
        res = 0.0
        for i in range(n):
            res += float(a[i])
        return res
    if n <= 128:
        r = [float(a[j]) for j in range(8)]
        i = 8
        while i < n - (n % 8):
            for j in range(8):
                r[j] += float(a[i + j])
            i += 8
        res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]))
        while i < n:
            res += float(a[i])
            i += 1
        return res
    n2 = n // 2
    n2 -= n2 % 8
    return numpy_pairwise_sum(a[:n2]) + numpy_pairwise_sum(a[n2:])


# --------------------------------------------------------------------------------------------
# Convergence test and maxima over the value function (row a8)
# --------------------------------------------------------------------------------------------

def max_values(beliefs: np.ndarray, alphas: np.ndarray):
    """(max_v b.alpha_v, first-index argmax) -- src/pomdp.py:2165, 1639, 1735, 1393."""
    prod = np.matmul(beliefs, alphas.T)
    return np.max(prod, axis=1), np.argmax(prod, axis=1)


def compute_change(beliefs: np.ndarray, alphas_a: np.ndarray, alphas_b: np.ndarray) -> float:
    """max_b | max_v b.alpha_v - max_v' b.alpha'_v' | (src/pomdp.py:2165-2167)."""
    va = np.max(np.matmul(beliefs, alphas_a.T), axis=1)
    vb = np.max(np.matmul(beliefs, alphas_b.T), axis=1)
    return float(np.max(np.abs(vb - va)))


# --------------------------------------------------------------------------------------------
# MDP value iteration (row a11)
# --------------------------------------------------------------------------------------------

def vi_sweep(reach, probs, rbar, gamma, v_opt) -> np.ndarray:
    """alpha[a,s] = Rbar[s,a] + gamma * sum_r P[s,a,r] * V*[reach[s,a,r]] (src/mdp.py:1507)."""
    return rbar.T + (gamma * np.einsum('sar,sar->as', probs, v_opt[reach]))


def vi_solve(reach, probs, rbar, gamma: float, eps: float, horizon: int = 10000):
    """
    Value iteration (src/mdp.py:1485-1525).  Each sweep's alpha set is byte-deduped exactly like the
    ValueFunction constructor (first position, last action).  Returns (alphas, actions, iterations).
    """
    A = reach.shape[1]
    alphas, actions, _ = dedup_rows(np.ascontiguousarray(rbar.T), np.arange(A))
    v_opt = np.max(alphas, axis=0)
    limit = eps * (gamma / (1 - gamma))
    it = 0
    for it in range(1, horizon + 1):
        old = v_opt
        alphas, actions, _ = dedup_rows(vi_sweep(reach, probs, rbar, gamma, v_opt), np.arange(A))
        v_opt = np.max(alphas, axis=0)
        if np.max(np.abs(v_opt - old)) < limit:
            break
    return alphas, actions, it


# --------------------------------------------------------------------------------------------
# Pruning (src/mdp.py:857-866) and HSVI's sawtooth upper bound (src/pomdp.py:873-895)
# --------------------------------------------------------------------------------------------

def prune_pointwise_dominated(alphas: np.ndarray) -> np.ndarray:
    """Indices kept by prune level 2: v survives iff exactly one vector (itself) is >= v everywhere."""
    keep = []
    for i, v in enumerate(alphas):
        if np.sum(np.all(alphas >= v, axis=1)) == 1:
            keep.append(i)
    return np.array(keep, dtype=np.int64)


def sawtooth_reference(corner_values, ub_beliefs, ub_values, belief) -> float:
    """The reference's formula verbatim (src/pomdp.py:887-895): min over ALL coordinates, NaN where 0/0."""
    v0 = np.dot(belief, corner_values)
    if ub_beliefs.shape[0] == 0:
        return float(v0)
    with np.errstate(divide='ignore', invalid='ignore'):
        vb = v0 + ((ub_values - np.dot(ub_beliefs, corner_values)) * np.min(belief / ub_beliefs, axis=1))
    return float(np.min(np.append(vb, v0)))


def sawtooth_intended(corner_values, ub_beliefs, ub_values, belief) -> float:
    """
    The intended sawtooth (Shani et al.): ratio minimised over coordinates with b_i[s] > 0 only.  Equal to
    `sawtooth_reference` whenever the stored beliefs are strictly positive (e.g. tiger); on sparse
    beliefs the reference yields NaN and the engine follows this definition instead.
    """
    v0 = np.dot(belief, corner_values)
    if ub_beliefs.shape[0] == 0:
        return float(v0)
    with np.errstate(divide='ignore', invalid='ignore'):
        ratio = np.where(ub_beliefs > 0, belief[None, :] / ub_beliefs, np.inf)
    vb = v0 + (ub_values - np.dot(ub_beliefs, corner_values)) * np.min(ratio, axis=1)
    return float(np.min(np.append(vb, v0)))


# --------------------------------------------------------------------------------------------
# Deterministic expansion scorings (row a9), ingredient level
# --------------------------------------------------------------------------------------------

def ssea_min_distances(beliefs: np.ndarray, successors: np.ndarray) -> np.ndarray:
    """dist[n,a,o] = min_b || beliefs[b] - successors[n,a,o] ||_2 (src/pomdp.py:1682-1686)."""
    diff = beliefs[:, None, None, None, :] - successors
    dist = np.sqrt(np.einsum('bnaos,bnaos->bnao', diff, diff))
    return np.min(dist, axis=0)


def ger_scores(beliefs, successors, alphas, rto, gamma, r_min_raw, r_max_raw):
    """
    GER error terms (src/pomdp.py:1728-1754): returns (res[b,a], bao_probs[b,a,o], eps[b,a,o]).
    """
    r_min = r_min_raw / (1 - gamma)
    r_max = r_max_raw / (1 - gamma)
    best = np.argmax(np.dot(beliefs, alphas.T), axis=1)
    b_alphas = alphas[best]
    b_diffs = successors - beliefs[:, None, None, :]
    alphas_p = np.where(b_diffs >= 0, r_max, r_min)
    alphas_diffs = alphas_p - b_alphas[:, None, None, :]
    eps = np.einsum('baos,baos->bao', alphas_diffs, b_diffs)
    bao_probs = np.einsum('bs,saor->bao', beliefs, rto)
    res = np.einsum('bao,bao->ba', bao_probs, eps)
    return res, bao_probs, eps
