"""
The sharded solve loop (`PBVI_Solver.solve(..., group=...)`, parallel.ShardedSolveState) on the REAL engine: two ranks share the one
GPU of the test box (gloo group; device tensors are staged through the host for the collectives -- NCCL refuses two ranks per
device) and must return the value function and the history counts a single process computes with the same seeds: expansion on rank 0
+ broadcast, append-only ownership, position-carrying tuple exchange, sharded compute_change.
`tools/check_sharded.py` runs the same comparison over NCCL on real multi-GPU boxes.
"""
import os
import random
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


CASES = {
    # name: (solver factory args, solve kwargs, replicate_below)
    'fsvi_new_points_replicated': (('fsvi',), dict(expansions=10, max_belief_growth=20), 64),
    'fsvi_new_points_sharded': (('fsvi',), dict(expansions=8, max_belief_growth=30), 1),
    'perseus_full_backup': (('perseus',), dict(expansions=6, max_belief_growth=40, full_backup=True), 64),
    'ssra_full_backup': (('ssra',), dict(expansions=5, max_belief_growth=16), 64),
    'hsvi': (('hsvi',), dict(expansions=4, max_belief_growth=12), 64),
    'perseus_limited': (('perseus',), dict(expansions=8, max_belief_growth=25, limit_value_function_size=30), 64),
}


def _solve(case, group):
    import torch
    from pomdp_pbvi_exploration_b200 import FSVI_Solver, HSVI_Solver, PBVI_Solver
    from pomdp_pbvi_exploration_b200.recipes import olfactory_wrap_model
    (flavour,), kw, replicate_below = CASES[case]
    model = olfactory_wrap_model(points_per_unit=6)
    np.random.seed(11)
    random.seed(11)
    if flavour == 'fsvi':
        solver = FSVI_Solver(gamma=0.99, eps=1e-6)
    elif flavour == 'hsvi':
        solver = HSVI_Solver(gamma=0.99, eps=1e-6)
    else:
        solver = PBVI_Solver(gamma=0.99, eps=1e-6, expand_function=flavour)
    extra = dict(group=group, replicate_below=replicate_below) if group is not None else {}
    vf, hist = solver.solve(model, print_progress=False, **kw, **extra)
    torch.cuda.synchronize()
    stats = dict(solver._shard_state.stats) if solver._shard_state is not None else {}
    return (vf.alpha_vector_array.cpu().numpy(), vf.actions.copy(), list(hist.alpha_vector_counts), list(hist.beliefs_counts),
            [float(x) for x in hist.value_function_changes], stats)


def _worker(rank, world, port, case, queue):
    import torch
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(0)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    if rank != 0:
        np.random.seed(999 + rank)            # the other ranks' host RNG must not matter: every draw happens on rank 0
        random.seed(999 + rank)
    out = _solve(case, True)
    queue.put((rank,) + out)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize('case', list(CASES))
def test_sharded_solve_equals_single_process(case):
    import torch
    import torch.multiprocessing as mp
    assert torch.cuda.is_available()
    want = _solve(case, None)
    world = 2
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, case, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=600) for _ in range(world)]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for rank, rows, actions, n_alpha, n_belief, changes, stats in results:
        assert np.array_equal(rows, want[0]), (case, rank)
        assert np.array_equal(actions, want[1]), (case, rank)
        assert n_alpha == want[2] and n_belief == want[3], (case, rank)
        assert changes == want[4], (case, rank)
        if CASES[case][1].get('full_backup') or case in ('ssra_full_backup', 'fsvi_new_points_sharded'):
            assert stats['sharded_backups'] > 0
        if case == 'fsvi_new_points_replicated':
            assert stats['replicated_backups'] > 0 and stats['sharded_backups'] == 0
