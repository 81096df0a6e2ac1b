"""Host packer throughput (pbvi_pack_rows_host) vs thread count on the bench beliefs, next to the plain pinned H2D copy."""
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from pomdp_pbvi_exploration_b200.recipes import olfactory_wrap_model  # noqa: E402


def main():
    model = olfactory_wrap_model()
    dev = model.device
    _, beliefs, _, _ = bench.build_workload(model, 10000, 40, seed=0)
    host = beliefs.cpu().pin_memory()
    nB, S = host.shape
    n_c, W = dev.pack_geometry(S)
    SL = 256
    n_slabs = -(-nB // SL)
    region = SL * n_c * 4 + 4
    bm = torch.empty((nB, W), dtype=torch.int32).pin_memory()
    rs = torch.empty((n_slabs, SL + 1), dtype=torch.int32).pin_memory()
    pk = torch.empty((n_slabs * region,), dtype=torch.float64).pin_memory()

    def pack(i):
        r0, r1 = i * SL, min(nB, (i + 1) * SL)
        return dev.pack_rows_host(host[r0:r1], bm[r0:r1], rs[i], pk[i * region:(i + 1) * region])

    for threads in (1, 2, 4, 8, 16, 32):
        pool = ThreadPoolExecutor(max_workers=threads)
        best = 1e9
        for _ in range(3):
            t0 = time.perf_counter()
            tot = sum(pool.map(pack, range(n_slabs)))
            best = min(best, time.perf_counter() - t0)
        print(f'{threads:2d} threads: {best * 1e3:7.2f} ms  {nB * S * 8 / best / 1e9:6.1f} GB/s scanned, packed share {tot * 32 / (nB * S * 8):.3f}')
    d = torch.empty_like(host, device=dev.device)
    for _ in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        d.copy_(host, non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    print(f'plain pinned H2D: {dt * 1e3:7.2f} ms  {nB * S * 8 / dt / 1e9:6.1f} GB/s')


if __name__ == '__main__':
    main()
