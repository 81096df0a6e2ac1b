"""cProfile of a short solve (which host calls dominate the expand step?).  python tools/profile_expand.py fsvi 40 100"""
import cProfile
import os
import pstats
import random
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pomdp_pbvi_exploration_b200 import FSVI_Solver, HSVI_Solver, PBVI_Solver  # noqa: E402
from pomdp_pbvi_exploration_b200.recipes import olfactory_wrap_model  # noqa: E402

flavour, expansions, growth = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
model = olfactory_wrap_model()
model.device
np.random.seed(0); random.seed(0)
solver = {'fsvi': FSVI_Solver, 'hsvi': HSVI_Solver}.get(flavour, lambda **k: PBVI_Solver(expand_function=flavour, **k))(gamma=0.99, eps=1e-6)
solver.solve(model, expansions=3, max_belief_growth=growth, print_progress=False)      # warm-up (VI, allocations)
pr = cProfile.Profile()
pr.enable()
vf, hist = solver.solve(model, expansions=expansions, max_belief_growth=growth, print_progress=False)
torch.cuda.synchronize()
pr.disable()
print('expand avg %.4f s, backup avg %.4f s, |V|=%d |B|=%d' % (np.mean(hist.expansion_times), np.mean(hist.backup_times), len(vf), hist.beliefs_counts[-1]))
pstats.Stats(pr).sort_stats('cumulative').print_stats(32)
