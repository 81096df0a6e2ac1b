"""
Full solves of the 22021-state olfactory POMDP in the as-published shapes:
    FSVI, 300 expansions x 100 beliefs, gamma 0.99, eps 1e-6   (reference: 2983.5 s NumPy CPU, 204.9 s CuPy GPU)
    python tools/solve_olfactory.py [flavour] [expansions] [growth] [full] [update_passes]
Prints the reference-format summary plus wall time split into expand / backup / change / other.
"""
import json
import os
import random
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pomdp_pbvi_exploration_b200 import FSVI_Solver, HSVI_Solver, PBVI_Solver  # noqa: E402
from pomdp_pbvi_exploration_b200.recipes import olfactory_wrap_model  # noqa: E402


def main():
    flavour = sys.argv[1] if len(sys.argv) > 1 else 'fsvi'
    expansions = int(sys.argv[2]) if len(sys.argv) > 2 else 300
    growth = int(sys.argv[3]) if len(sys.argv) > 3 else 100
    model = olfactory_wrap_model()
    model.device                                   # create the handle outside the timed region (the reference times solve() only)
    np.random.seed(0)
    random.seed(0)
    full = len(sys.argv) > 4 and sys.argv[4] == 'full'
    if flavour == 'fsvi' and not full:
        solver = FSVI_Solver(gamma=0.99, eps=1e-6)
    elif flavour == 'hsvi':
        solver = HSVI_Solver(gamma=0.99, eps=1e-6)
    else:
        solver = PBVI_Solver(gamma=0.99, eps=1e-6, expand_function=flavour)
    change_s = [0.0]
    orig_change = solver.compute_change

    def timed_change(*a, **k):
        torch.cuda.synchronize()
        t = time.perf_counter()
        out = orig_change(*a, **k)
        torch.cuda.synchronize()
        change_s[0] += time.perf_counter() - t
        return out
    solver.compute_change = timed_change
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    kw = dict(full_backup=True) if len(sys.argv) > 4 and sys.argv[4] == 'full' else {}
    if len(sys.argv) > 5:
        kw['update_passes'] = int(sys.argv[5])
    if os.environ.get('PBVI_CPROFILE'):
        import cProfile, pstats, io
        pr = cProfile.Profile(); pr.enable()
    vf, hist = solver.solve(model, expansions=expansions, max_belief_growth=growth, print_progress=False, **kw)
    if os.environ.get('PBVI_CPROFILE'):
        pr.disable(); st = io.StringIO(); pstats.Stats(pr, stream=st).sort_stats('tottime').print_stats(22); print(st.getvalue()[:5000])
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    print(hist.summary)
    pairs = sum(b * v for b, v in zip(np.diff(hist.beliefs_counts) if not hist.expand_append else hist.beliefs_counts[1:], hist.alpha_vector_counts[:-1])) if 'update_passes' not in kw else 0
    out = dict(flavour=flavour, expansions=len(hist.expansion_times), growth=growth, wall_s=wall, expand_s=sum(hist.expansion_times),
               backup_s=sum(hist.backup_times), change_s=change_s[0], final_alphas=len(vf), final_beliefs=hist.beliefs_counts[-1],
               backup_pairs=float(pairs), backup_pairs_per_s=float(pairs) / max(sum(hist.backup_times), 1e-9),
               backup_ms_at=[round(1e3 * hist.backup_times[i], 2) for i in range(0, len(hist.backup_times), max(1, len(hist.backup_times) // 10))],
               expand_ms_at=[round(1e3 * hist.expansion_times[i], 2) for i in range(0, len(hist.expansion_times), max(1, len(hist.expansion_times) // 10))],
               reference_published={'cpu_s': 2983.5, 'cupy_gpu_s': 204.9, 'source': 'Olfactory_Alternation_Paper_Wrap.ipynb[43],[30]'})
    print(json.dumps(out))


if __name__ == '__main__':
    main()
