"""Times the phases of the host-buffer (e2e) backup step of bench.py on one GPU (late value function): where do the milliseconds go?"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from pomdp_pbvi_exploration_b200 import BeliefSet, ValueFunction  # noqa: E402
from pomdp_pbvi_exploration_b200.recipes import olfactory_wrap_model  # noqa: E402


def main():
    model = olfactory_wrap_model()
    dev = model.device
    nB = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
    solver, beliefs, vfs, _ = bench.build_workload(model, nB, 1000, seed=0)
    vf = vfs['late']
    h_b = beliefs.cpu().pin_memory()
    h_a = vf.alpha_vector_array.cpu().pin_memory()
    acts = vf.actions.copy()
    rows = []
    for it in range(5):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        bs = BeliefSet(model, h_b)
        t1 = time.perf_counter()
        v_in = ValueFunction(model, h_a, acts)
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        job = bs.__dict__.get('_pack_job')
        res = solver.backup(model, bs, v_in, append=False, belief_dominance_prune=False)
        torch.cuda.synchronize()
        t3 = time.perf_counter()
        out = res.numpy(staged=True)
        torch.cuda.synchronize()
        t4 = time.perf_counter()
        pack_done = None
        if job is not None:
            for f in job.futures:
                f.result()
        rows.append([(t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, (t4 - t3) * 1e3, (t4 - t0) * 1e3])
    r = np.array(rows[1:]).mean(0)
    print(f'BeliefSet(host) ctor (starts the packers) {r[0]:7.2f} ms\\nValueFunction(host) ctor (H2D 176 MB + dedup) {r[1]:7.2f} ms\\n'
          f'solver.backup (packed upload streamed behind select, assemble, dedup) {r[2]:7.2f} ms\\nValueFunction.numpy(staged) D2H {r[3]:7.2f} ms\\n'
          f'step {r[4]:7.2f} ms; h2d bytes {solver.last_h2d_bytes / 1e6:.0f} MB; new rows {len(res)}')
    # device-resident reference and the packers alone
    bs_d = BeliefSet(model, beliefs)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    solver.backup(model, bs_d, vf, append=False, belief_dominance_prune=False)
    torch.cuda.synchronize(); print(f'device-resident backup {(time.perf_counter() - t0) * 1e3:7.2f} ms')
    t0 = time.perf_counter()
    job = dev.start_pack(h_b)
    for f in job.futures:
        f.result()
    job.close()
    print(f'packers alone ({len(job.futures)} threads) {(time.perf_counter() - t0) * 1e3:7.2f} ms')
    t0 = time.perf_counter(); x = h_b.cuda(non_blocking=True); torch.cuda.synchronize()
    print(f'plain pinned H2D of the dense beliefs {(time.perf_counter() - t0) * 1e3:7.2f} ms ({h_b.numel() * 8 / 1e9 / (time.perf_counter() - t0):.1f} GB/s)')


if __name__ == '__main__':
    main()
