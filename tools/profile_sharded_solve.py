"""
torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/profile_sharded_solve.py [expansions]
Where does a sharded FSVI solve spend its host time?  cProfile of rank 0 (top cumulative entries) + a micro-timing of the row broadcast.
"""
import cProfile
import io
import os
import pstats
import random
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pomdp_pbvi_exploration_b200 import FSVI_Solver  # noqa: E402
from pomdp_pbvi_exploration_b200.parallel import broadcast_  # noqa: E402
from pomdp_pbvi_exploration_b200.recipes import olfactory_wrap_model  # noqa: E402


def main():
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    n_exp = int(sys.argv[1]) if len(sys.argv) > 1 else 60
    model = olfactory_wrap_model()
    rows = torch.rand((100, model.state_count), dtype=torch.float64, device='cuda')
    head = torch.zeros((1,), dtype=torch.int64, device='cuda')
    for _ in range(5):
        broadcast_(head, 0); broadcast_(rows, 0)
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    for _ in range(100):
        broadcast_(head, 0)
        n = int(head[0])
        broadcast_(rows, 0)
    torch.cuda.synchronize()
    if rank == 0:
        print(f'broadcast of a count + 100 x {model.state_count} rows: {(time.perf_counter() - t0) * 10:.3f} ms per pair of calls', flush=True)
    for grp in (None, True):
        np.random.seed(0); random.seed(0)
        solver = FSVI_Solver(gamma=0.99, eps=1e-6)
        pr = cProfile.Profile()
        torch.cuda.synchronize(); dist.barrier()
        t0 = time.perf_counter()
        pr.enable()
        vf, hist = solver.solve(model, expansions=n_exp, max_belief_growth=100, print_progress=False, **({'group': True} if grp else {}))
        pr.disable()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        if rank == 0:
            print(f'--- {"sharded" if grp else "single"}: wall {wall:.3f}s expand {sum(hist.expansion_times):.3f} backup {sum(hist.backup_times):.3f}', flush=True)
            st = io.StringIO()
            pstats.Stats(pr, stream=st).sort_stats('cumulative').print_stats(28)
            print(st.getvalue()[:6000], flush=True)
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
