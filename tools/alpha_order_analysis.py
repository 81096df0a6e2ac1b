"""
How much alpha-side work would a better COLUMN ORDER of the value function save?  The score kernel skips a (4-state chunk, 64-alpha column
quarter) when no alpha of the quarter is non-zero on the states the chunk lands on; which alphas share a quarter is just their order.
Proxy: live (chunk, quarter) cells of the alpha pattern itself (identity landing), for several orderings of the late value function.
    python tools/alpha_order_analysis.py
Result (round 2): on the late value function the proxy drops from 42.9 % (identity) to 30.1 % (sorted by the first non-zero state; a greedy
union clustering: 28.6 %) -- but the flops the score kernel actually executes only fell by 6 % when the select path sorted its columns that
way (7.43e11 -> 6.98e11; kernel 22.5 -> 21.7 ms), because most of the alpha cells that disappear lie where the beliefs are zero anyway, and
the two extra passes over the alphas cost 0.5 ms per call (young value function: 3.5 -> 4.1 ms per step).  The re-ordering was dropped.
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from pomdp_pbvi_exploration_b200.recipes import olfactory_wrap_model  # noqa: E402


def live_cells(nz_chunks: torch.Tensor, order: torch.Tensor, q: int = 64) -> int:
    x = nz_chunks[order]
    V = x.shape[0]
    pad = (-V) % q
    if pad:
        x = torch.cat([x, torch.zeros((pad, x.shape[1]), dtype=torch.bool, device=x.device)])
    return int(x.reshape(-1, q, x.shape[1]).any(dim=1).sum())


def main():
    model = olfactory_wrap_model()
    _, beliefs, vfs, info = bench.build_workload(model, 2000, 1000, seed=0)
    for name in ('late', 'young'):
        A = vfs[name].alpha_vector_array
        V, S = A.shape
        nz = A != 0
        Sp = (S + 3) // 4 * 4
        nzp = torch.zeros((V, Sp), dtype=torch.bool, device=A.device)
        nzp[:, :S] = nz
        chunks = nzp.reshape(V, Sp // 4, 4).any(dim=2)
        ident = torch.arange(V, device=A.device)
        idx = torch.arange(S, device=A.device, dtype=torch.float64)
        cnt = nz.sum(1).double().clamp(min=1)
        centroid = (nz.double() * idx).sum(1) / cnt
        first = torch.where(nz.any(1), nz.double().argmax(1), torch.full((V,), S, device=A.device))
        W = 361
        row_c = (nz.double() * (idx // W)).sum(1) / cnt
        col_c = (nz.double() * (idx % W)).sum(1) / cnt
        res = {'identity': live_cells(chunks, ident), 'by nnz': live_cells(chunks, torch.argsort(cnt)),
               'by centroid': live_cells(chunks, torch.argsort(centroid)), 'by first nz': live_cells(chunks, torch.argsort(first)),
               'by column centroid': live_cells(chunks, torch.argsort(col_c)), 'by (nnz bucket, col centroid)': live_cells(chunks, torch.argsort((cnt.log2().floor() * 1000 + col_c))),
               'random': live_cells(chunks, torch.randperm(V, device=A.device))}
        # greedy clustering: repeatedly start a quarter with the sparsest unassigned alpha and add the 63 alphas that enlarge the union least
        rem = list(range(V))
        order = []
        cf = chunks.float()
        while rem:
            r = torch.tensor(rem, device=A.device)
            seed = r[torch.argmin(cnt[r])]
            union = chunks[seed].clone()
            group = [int(seed)]
            rem.remove(int(seed))
            while len(group) < 64 and rem:
                r = torch.tensor(rem, device=A.device)
                extra = (cf[r] * (~union).float()).sum(1)
                pick = int(r[torch.argmin(extra)])
                group.append(pick)
                rem.remove(pick)
                union |= chunks[pick]
            order += group
        res['greedy union'] = live_cells(chunks, torch.tensor(order, device=A.device))
        total = (V + 63) // 64 * chunks.shape[1]
        dens = float(nz.double().mean())
        print(name, f'alpha density {dens:.4f}; live (chunk, quarter) cells of {total}:', {k: f'{v} ({v / total:.3f})' for k, v in res.items()}, flush=True)


if __name__ == '__main__':
    main()
