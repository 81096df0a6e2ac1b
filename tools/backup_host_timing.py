"""pbvi_backup_host (the C ABI's host-buffer backup: one alpha row per belief, no dedup) on the bench workload with pinned buffers:
time per call of the two-deep chunk pipeline next to its three parts run one after the other (upload, pbvi_backup, download)."""
import ctypes
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from pomdp_pbvi_exploration_b200.recipes import olfactory_wrap_model  # noqa: E402


def main():
    model = olfactory_wrap_model()
    dev = model.device
    nB = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
    solver, beliefs, vfs, _ = bench.build_workload(model, nB, 1000, seed=0)
    V = vfs['late'].alpha_vector_array
    hb, hv = beliefs.cpu().pin_memory(), V.cpu().pin_memory()
    out = torch.empty((nB, dev.S), dtype=torch.float64).pin_memory()
    act = torch.empty((nB,), dtype=torch.int32).pin_memory()
    st = torch.cuda.current_stream().cuda_stream
    ts = []
    for _ in range(5):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        rc = dev._lib.pbvi_backup_host(dev._h, hb.data_ptr(), nB, hv.data_ptr(), V.shape[0], ctypes.c_double(0.99), out.data_ptr(), act.data_ptr(), st)
        ts.append((time.perf_counter() - t0) * 1e3)
        assert rc == 0
    print(f'pbvi_backup_host, {nB} x {V.shape[0]}, pinned buffers: {[round(t, 1) for t in ts]} ms per call')
    # the three parts one after the other
    ts = []
    out2 = torch.empty_like(out).pin_memory()
    for _ in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        db, dv = hb.cuda(non_blocking=True), hv.cuda(non_blocking=True)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        rows, a, _, _ = dev.backup(db, dv, 0.99)
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        out2.copy_(rows, non_blocking=True)
        torch.cuda.synchronize()
        t3 = time.perf_counter()
        ts.append([(t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, (t3 - t0) * 1e3])
    m = np.array(ts[1:]).mean(0)
    print(f'one after the other: upload {m[0]:.1f} + pbvi_backup {m[1]:.1f} + download {m[2]:.1f} = {m[3]:.1f} ms')
    print('rows equal:', bool(torch.equal(out, rows.cpu())), 'actions equal:', bool(torch.equal(act, a.cpu())))
    # the whole backup (dedup included) in one call: only the distinct rows come back
    n = ctypes.c_int()
    ts = []
    for _ in range(5):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        rc = dev._lib.pbvi_backup_host_unique(dev._h, hb.data_ptr(), nB, hv.data_ptr(), V.shape[0], ctypes.c_double(0.99), out.data_ptr(), nB,
                                              act.data_ptr(), ctypes.byref(n), st)
        ts.append((time.perf_counter() - t0) * 1e3)
        assert rc == 0
    pb, pv = beliefs.cpu().numpy(), V.cpu().numpy()            # pageable arrays: what a NumPy host holds
    po, pa = np.empty((nB, dev.S)), np.empty(nB, dtype=np.int32)
    tp = []
    for _ in range(4):
        t0 = time.perf_counter()
        rc = dev._lib.pbvi_backup_host_unique(dev._h, pb.ctypes.data, nB, pv.ctypes.data, V.shape[0], ctypes.c_double(0.99), po.ctypes.data, nB,
                                              pa.ctypes.data, ctypes.byref(n), st)
        tp.append((time.perf_counter() - t0) * 1e3)
        assert rc == 0
    print(f'pbvi_backup_host_unique from pageable NumPy arrays: {[round(t, 1) for t in tp]} ms per call')
    from pomdp_pbvi_exploration_b200 import BeliefSet
    want = solver.backup(model, BeliefSet(model, beliefs), vfs['late'], append=False, belief_dominance_prune=False)
    wr, wa = want.numpy()
    print(f'pbvi_backup_host_unique: {[round(t, 1) for t in ts]} ms per call, {n.value} rows; equal to PBVI_Solver.backup: '
          f'{bool(np.array_equal(out[:n.value].numpy(), wr) and np.array_equal(act[:n.value].numpy(), wa))}')


if __name__ == '__main__':
    main()
