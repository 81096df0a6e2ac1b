"""A/B of the grouped assemble kernel's (tuples per block, states per thread): python tools/assemble_ab.py  (PBVI_B200_LIB selects the build)"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from pomdp_pbvi_exploration_b200.recipes import olfactory_wrap_model  # noqa: E402


def main():
    model = olfactory_wrap_model()
    dev = model.device
    solver, beliefs, vfs, _ = bench.build_workload(model, 10000, 1000, seed=0)
    vf = vfs['late']
    from pomdp_pbvi_exploration_b200 import BeliefSet
    tuples, first, last = solver.select_tuples_device(model, BeliefSet(model, beliefs), vf)
    V = vf.alpha_vector_array
    for rep in (1, 8):                                   # the local set, and an 8-rank merged set's size
        t = tuples.repeat(rep, 1)
        for _ in range(3):
            rows, keys = dev.backup_assemble(V, 0.99, t[:, 0], t[:, 1:], with_hash=True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            rows, keys = dev.backup_assemble(V, 0.99, t[:, 0], t[:, 1:], with_hash=True)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        ref = dev.row_hash(rows)
        print(f'{os.path.basename(os.environ.get("PBVI_B200_LIB", "default"))}: {t.shape[0]} tuples: {ms * 1e3:8.1f} us per call (incl. scan + order + finalise), '
              f'{t.shape[0] * dev.S * 8 / ms / 1e9:.2f} TB/s written; keys ok {bool(torch.equal(ref, keys))}', flush=True)


if __name__ == '__main__':
    main()
