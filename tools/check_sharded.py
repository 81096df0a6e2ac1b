"""
torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/check_sharded.py
Determinism check of the sharded backup on real GPUs: every rank's merged value function must equal, bit for bit and in
order, the value function a single process computes over the whole belief set (rank 0 recomputes it locally).
"""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pomdp_pbvi_exploration_b200 import BeliefSet, PBVI_Solver, ShardedBackup, ValueFunction  # noqa: E402
from pomdp_pbvi_exploration_b200.recipes import olfactory_wrap_model, perseus_walk_beliefs  # noqa: E402


def main():
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    model = olfactory_wrap_model()
    native = None
    if '--native' in sys.argv:
        # the library's own NCCL binding (pbvi_comm_*): the id travels over the torch group, every collective of the checks below
        # then goes through the C ABI
        from pomdp_pbvi_exploration_b200 import NativeComm
        idt = torch.zeros((128,), dtype=torch.uint8, device='cuda')
        if rank == 0:
            idt = torch.tensor(list(NativeComm.unique_id()), dtype=torch.uint8, device='cuda')
        dist.broadcast(idt, 0)
        native = NativeComm(model, rank, world, bytes(idt.cpu().tolist()))
        if rank == 0:
            print('collectives: libpbvi_b200 NCCL binding (pbvi_comm_init / pbvi_allgather_tuples / pbvi_allreduce_max / pbvi_broadcast_rows)', flush=True)
    B = perseus_walk_beliefs(model, 1500, seed=5)                      # same beliefs on every rank (same seed)
    g = dict(np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden', 'backup_olfactory_wrap.npz')))
    vf = ValueFunction(model, g['alphas'], g['alpha_actions'])
    solver = PBVI_Solver(gamma=0.99, eps=1e-6, expand_function='perseus')
    for mode, append in [('tuples', False), ('tuples', True), ('rows', False), ('rows', True)]:
        sb = ShardedBackup(solver, model, group=native, exchange=mode)
        lo, hi = sb.bounds(B.shape[0])
        merged = sb.backup(BeliefSet(model, B[lo:hi]), vf, append=append)
        rows, actions = merged.numpy()
        single = solver.backup(model, BeliefSet(model, B), vf, append=append, belief_dominance_prune=False)
        srows, sactions = single.numpy()
        ok = rows.shape == srows.shape and np.array_equal(rows, srows) and np.array_equal(actions, sactions)
        flag = torch.tensor([int(ok)], device='cuda')
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if rank == 0:
            print(f'exchange={mode} append={append}: world={world} merged {rows.shape[0]} rows, single-process {srows.shape[0]} rows, identical on all ranks: {bool(flag[0])}',
                  flush=True)
        assert bool(flag[0])
        chg = sb.compute_change(vf, merged, BeliefSet(model, B[lo:hi]))
        ref = solver.compute_change(vf, single, BeliefSet(model, B))
        assert chg == ref, (chg, ref)
    # ---- the sharded solve loop: same value function, same history counts as one process (every rank checks against its own
    #      single-process run with the same seeds; only rank 0's host RNG matters in the sharded run)
    import random
    from pomdp_pbvi_exploration_b200 import FSVI_Solver
    cases = [('fsvi 40x100 (new-points backups)', lambda: FSVI_Solver(gamma=0.99, eps=1e-6), dict(expansions=40, max_belief_growth=100)),
             ('perseus full backup 3x1500', lambda: PBVI_Solver(gamma=0.99, eps=1e-6, expand_function='perseus'),
              dict(expansions=3, max_belief_growth=1500, full_backup=True))]
    for name, make, kw in cases:
        outs = []
        for grp in (None, True):
            np.random.seed(5 if (grp is None or rank == 0) else 1000 + rank)
            random.seed(5 if (grp is None or rank == 0) else 1000 + rank)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            vf_s, hist = make().solve(model, print_progress=False, **kw, **({'group': native if native is not None else True} if grp else {}))
            torch.cuda.synchronize()
            outs.append((vf_s.numpy(), hist.alpha_vector_counts, hist.beliefs_counts, hist.value_function_changes, time.perf_counter() - t0))
        (r1, a1), (r2, a2) = outs[0][0], outs[1][0]
        ok = r1.shape == r2.shape and np.array_equal(r1, r2) and np.array_equal(a1, a2) and outs[0][1:4] == outs[1][1:4]
        flag = torch.tensor([int(ok)], device='cuda')
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if rank == 0:
            print(f'solve {name}: world={world} |V|={r2.shape[0]} |B|={outs[1][2][-1]} single {outs[0][4]:.2f}s sharded {outs[1][4]:.2f}s '
                  f'identical on all ranks: {bool(flag[0])}', flush=True)
        assert bool(flag[0])
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
