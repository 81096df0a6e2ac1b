"""Times the phases of one device-resident `PBVI_Solver.backup` step of bench.py on one GPU (synchronised wall times)."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from pomdp_pbvi_exploration_b200 import BeliefSet  # noqa: E402
from pomdp_pbvi_exploration_b200.recipes import olfactory_wrap_model  # noqa: E402


def timed(fn, n=5):
    out, ts = None, []
    for _ in range(n):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = fn()
        torch.cuda.synchronize()
        ts.append((time.perf_counter() - t0) * 1e3)
    return out, min(ts)


def main():
    model = olfactory_wrap_model()
    dev = model.device
    nB = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
    solver, beliefs, vfs, _ = bench.build_workload(model, nB, 1000, seed=0)
    vf = vfs[os.environ.get('PBVI_VF', 'late')]
    bs = BeliefSet(model, beliefs)
    V = vf.alpha_vector_array
    _, t = timed(lambda: dev.backup_select(beliefs, V, 0.99)); print(f'backup_select (kernels)      {t:8.3f} ms')
    (tuples, first, last), t = timed(lambda: solver.select_tuples(model, bs, vf)); print(f'select_tuples (incl. host)   {t:8.3f} ms  -> {tuples.shape[0]} tuples')
    (rows, keys), t = timed(lambda: dev.backup_assemble(V, 0.99, tuples[:, 0], tuples[:, 1:], with_hash=True)); print(f'backup_assemble + keys       {t:8.3f} ms')
    out, t = timed(lambda: solver.rows_from_tuples(model, vf, tuples, last)); print(f'rows_from_tuples             {t:8.3f} ms  -> {len(out)} rows')
    out, t = timed(lambda: solver.backup(model, bs, vf, append=False, belief_dominance_prune=False)); print(f'solver.backup                {t:8.3f} ms')
    # the same at 8x the tuples (the merged set of an 8-rank sharded step)
    t8 = np.tile(tuples, (8, 1))[:6000]
    _, t = timed(lambda: dev.backup_assemble(V, 0.99, t8[:, 0], t8[:, 1:], with_hash=True)); print(f'backup_assemble x{t8.shape[0]}      {t:8.3f} ms')
    import cProfile
    import pstats
    pr = cProfile.Profile()
    pr.enable()
    solver.backup(model, bs, vf, append=False, belief_dominance_prune=False)
    torch.cuda.synchronize()
    pr.disable()
    pstats.Stats(pr).sort_stats('tottime').print_stats(14)


if __name__ == '__main__':
    main()
