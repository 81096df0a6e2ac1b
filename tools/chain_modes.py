"""Belief chains on the olfactory model: multi-launch (0) vs one persistent block (1) vs a cluster of 8 blocks (2); python tools/chain_modes.py"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pomdp_pbvi_exploration_b200.recipes import olfactory_wrap_model  # noqa: E402


def main():
    model = olfactory_wrap_model()
    dev = model.device
    rng = np.random.default_rng(0)
    n = 4000
    acts = rng.integers(0, 6, n).astype(np.int32)
    us = rng.random(n)
    b0 = torch.as_tensor(model.start_probabilities).cuda()
    ref = None
    for mode in (0, 1, 2, 1, 2):
        dev.set_option('chain_kernel', mode)
        dev.perseus_walk(b0, acts[:50], us[:50])
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out, obs = dev.perseus_walk(b0, acts, us, want_observations=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        same = True if ref is None else bool(torch.equal(out.nan_to_num(-1.0), ref.nan_to_num(-1.0)))
        if ref is None:
            ref = out
        print(f'mode {mode}: perseus walk of {n} steps {dt * 1e3:8.2f} ms = {dt / n * 1e6:6.2f} us/step; same bytes as mode 0: {same}', flush=True)
    dev.set_option('chain_kernel', 2)


if __name__ == '__main__':
    main()
