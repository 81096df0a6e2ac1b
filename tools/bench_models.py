"""
Backup throughput on the other BASELINE.json configurations (1 GPU), next to the CPU oracle where it fits:
  configs[0] tiger (S2 A3 O2 R2), configs[1] 4x4 grid dense (S16 A4 O2 R15) at B x V in {16,256,4096} x {4,64,1024},
  configs[4] synthetic sparse sweep S in {1k,10k,100k}, A/O/R varied, B=1024 k-sparse beliefs, V=256.
Prints one JSON object per point.  `pairs_per_s` = B*V / CUDA-event time of PBVI_Solver.backup (select + assemble + dedup).
    python tools/bench_models.py [--cpu] [--quick]
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pomdp_pbvi_exploration_b200 import BeliefSet, Model, PBVI_Solver, ValueFunction  # noqa: E402
from pomdp_pbvi_exploration_b200.recipes import synthetic_sparse_model, tiger_model  # noqa: E402

CPU = '--cpu' in sys.argv
QUICK = '--quick' in sys.argv


def golden_model(tag):
    m = dict(np.load(os.path.join(ROOT, 'tests', 'golden', f'model_{tag}.npz')))
    S, A, O = m['rto'].shape[0], m['rto'].shape[1], m['rto'].shape[2]
    return Model(states=S, actions=A, observations=O, transitions=m['transition_table'], rewards=m['reward_table'],
                 observation_table=m['obs_table'], start_probabilities=m['start']), float(m['gamma'])


def time_backup(model, gamma, B, V, acts, reps=5):
    solver = PBVI_Solver(gamma=gamma, eps=1e-6, expand_function='ssea')
    bs, vf = BeliefSet(model, B), ValueFunction(model, V, acts)
    for _ in range(2):
        out = solver.backup(model, bs, vf, append=False, belief_dominance_prune=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = solver.backup(model, bs, vf, append=False, belief_dominance_prune=False)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    stats = model.device.last_stats()
    return ms, len(out), len(vf), stats


def cpu_backup(model, gamma, B, V):
    from oracle import pbvi_oracle as orc
    t0 = time.perf_counter()
    out = orc.backup_chunked(model.reachable_states, model.reachable_transitional_observation_table, model.expected_rewards_table, gamma, B, V,
                             chunk=64)
    orc.dedup_rows(out['alpha'], out['a_star'])
    return time.perf_counter() - t0


def report(name, model, gamma, B, V, acts, cpu_ok):
    ms, n_out, n_v, stats = time_backup(model, gamma, B, V, acts)
    line = dict(config=name, S=model.state_count, A=model.action_count, O=model.observation_count, R=model.reachable_state_count,
                B=B.shape[0], V=n_v, ms_per_backup=round(ms, 4), pairs_per_s=B.shape[0] * n_v / (ms * 1e-3), new_alpha_rows=n_out,
                algorithmic_tflops=2.0 * model.action_count * model.observation_count * model.state_count * B.shape[0] * n_v / (ms * 1e-3) / 1e12)
    if CPU and cpu_ok:
        t = cpu_backup(model, gamma, B, V if isinstance(V, np.ndarray) else V.cpu().numpy())
        line['cpu_oracle_pairs_per_s'] = B.shape[0] * n_v / t
        line['speedup_vs_cpu_oracle'] = line['pairs_per_s'] / line['cpu_oracle_pairs_per_s']
    print(json.dumps(line), flush=True)


def dirichlet_beliefs(rng, n, S, k):
    B = np.zeros((n, S))
    for i in range(n):
        idx = rng.choice(S, min(k, S), replace=False)
        B[i, idx] = rng.dirichlet(np.ones(len(idx)))
    return B


def main():
    rng = np.random.default_rng(0)
    # ---- configs[0]: tiger
    model = tiger_model()
    B = dirichlet_beliefs(rng, 80, 2, 2)
    V = np.array([[-100.0, 10.0], [10.0, -100.0], [-1.0, -1.0], [3.0, 5.0], [5.0, 3.0], [-20.0, 8.0], [8.0, -20.0], [0.0, 0.0], [1.0, 2.0]])
    report('tiger', model, 0.95, B, V, rng.integers(0, 3, len(V)), True)
    # ---- configs[1]: 4x4 grid, dense transitions (R = 15)
    model, gamma = golden_model('grid4x4')
    for nB in ([16, 256] if QUICK else [16, 256, 4096]):
        for nV in ([4, 64] if QUICK else [4, 64, 1024]):
            B = np.concatenate([np.eye(16)[:min(16, nB)], dirichlet_beliefs(rng, max(0, nB - 16), 16, 16)])
            V = rng.random((nV, 16)) * 3
            report('grid4x4_R15', model, gamma, B, V, rng.integers(0, 4, nV), nB * nV <= 256 * 1024)
    # ---- configs[4]: synthetic sparse sweep
    sweep = [(1000, 4, 2, 1), (1000, 8, 4, 4), (10000, 4, 2, 1), (10000, 8, 4, 2), (10000, 16, 8, 8)]
    if not QUICK:
        sweep += [(30000, 8, 4, 1), (100000, 4, 2, 1), (100000, 16, 8, 4)]
    for S, A, O, R in sweep:
        model = synthetic_sparse_model(S, A, O, R, seed=1)
        B = dirichlet_beliefs(rng, 1024, S, 2048)
        V = rng.random((256, S))
        report('synthetic_sparse', model, 0.95, B, V, rng.integers(0, A, 256), S * A * O * 256 * 8 < 4e9)
        model.device.close()
        del model
        torch.cuda.empty_cache()


if __name__ == '__main__':
    main()
