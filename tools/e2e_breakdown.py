"""Times the phases of the host-buffer (e2e) backup path of bench.py on one GPU: where do the milliseconds go?"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from pomdp_pbvi_exploration_b200 import BeliefSet, ValueFunction  # noqa: E402
from pomdp_pbvi_exploration_b200.recipes import olfactory_wrap_model  # noqa: E402


def timed(fn, n=3):
    out = None
    ts = []
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
    solver, beliefs, vf, _ = bench.build_workload(model, nB, 1000, seed=0)
    h_b = beliefs.cpu().pin_memory()
    h_a = vf.alpha_vector_array.cpu().pin_memory()
    acts = vf.actions.copy()
    bs, t = timed(lambda: BeliefSet(model, h_b)); print(f'BeliefSet(host pinned)       {t:8.2f} ms  ({h_b.numel() * 8 / t / 1e6:.1f} GB/s)')
    v_in, t = timed(lambda: ValueFunction(model, h_a, acts)); print(f'ValueFunction(host pinned)   {t:8.2f} ms')
    _, t = timed(lambda: dev.backup_select(bs.belief_array, v_in.alpha_vector_array, 0.99)); print(f'backup_select                {t:8.2f} ms')
    out, t = timed(lambda: solver.backup(model, bs, v_in, append=False, belief_dominance_prune=False)); print(f'solver.backup (device)       {t:8.2f} ms  -> {len(out)} rows')
    _, t = timed(lambda: out.numpy()); print(f'ValueFunction.numpy() D2H    {t:8.2f} ms  ({len(out) * dev.S * 8 / t / 1e6:.1f} GB/s)')
    pin = torch.empty((len(out), dev.S), dtype=torch.float64).pin_memory()
    _, t = timed(lambda: pin.copy_(out.alpha_vector_array)); print(f'D2H into pinned buffer       {t:8.2f} ms')
    vstar, value, astar = dev.backup_select(bs.belief_array, v_in.alpha_vector_array, 0.99)
    torch.cuda.synchronize()
    import cProfile
    import pstats
    pr = cProfile.Profile()
    pr.enable()
    solver.backup(model, bs, v_in, append=False, belief_dominance_prune=False)
    torch.cuda.synchronize()
    pr.disable()
    pstats.Stats(pr).sort_stats('cumulative').print_stats(18)


if __name__ == '__main__':
    main()
