"""
How much of the score kernel's executed work is structurally necessary?  Counts, on the bench workload, the flops the
block-sparse GEMM executes at different skipping granules (rows per group x states per chunk x alphas per column block) and the
element-level lower bound (only non-zero belief x RTO x alpha products).  Runs on the GPU with torch ops (analysis only).

    python tools/sparsity_analysis.py [workload.pt]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from pomdp_pbvi_exploration_b200.recipes import olfactory_wrap_model  # noqa: E402


def main():
    model = olfactory_wrap_model()
    dev = model.device.device
    if len(sys.argv) > 1:
        blob = torch.load(sys.argv[1])
        beliefs, alphas = blob['beliefs'].to(dev), blob['alphas'].to(dev)
    else:
        _, beliefs, vfs, _ = bench.build_workload(model, 10000, 1000, seed=0)
        vf = vfs[os.environ.get('PBVI_VF', 'young')]           # 'young' (round 1's analysis) or 'late'
        alphas = vf.alpha_vector_array
    B, S = beliefs.shape
    V = alphas.shape[0]
    A, O = model.action_count, model.observation_count
    reach = torch.as_tensor(model.reachable_states[:, :, 0]).to(dev)                       # [S,A]
    rto = torch.as_tensor(model.reachable_transitional_observation_table[:, :, :, 0]).to(dev)   # [S,A,O]
    dense = 2.0 * B * V * A * O * S
    bnz = beliefs != 0
    anz = alphas != 0
    rnz = rto != 0
    print(f'B={B} V={V} S={S}; belief density {bnz.float().mean():.4f}, alpha density {anz.float().mean():.4f}, RTO density per o '
          f'{[round(float(rnz[:, :, o].float().mean()), 4) for o in range(O)]}')
    # element-level bound: sum_{b,s,a,o} [b!=0][rto!=0] * (#alphas non-zero at reach[s,a])
    cnt_alpha = anz.float().sum(0)                                                          # [S]
    per_state = torch.zeros(S, device=dev, dtype=torch.float64)
    for a in range(A):
        per_state += rnz[:, a, :].double().sum(1) * cnt_alpha[reach[:, a]].double()
    ideal = 2.0 * float((bnz.double().sum(0) * per_state).sum())
    print(f'element-level lower bound: {ideal:.4e} flops = {ideal / dense:.4%} of dense')

    def executed(rows, kc, cols):
        Sp = -(-S // kc) * kc
        nC = Sp // kc
        Bp = -(-B // rows) * rows
        bp = torch.zeros((Bp, Sp), dtype=torch.bool, device=dev)
        bp[:B, :S] = bnz
        G = bp.view(Bp // rows, rows, nC, kc).any(3).any(1).double().sum(0)                 # [nC] live row groups per chunk
        Vp = -(-V // cols) * cols
        ap = torch.zeros((Vp, S), dtype=torch.bool, device=dev)
        ap[:V] = anz
        alive = ap.view(Vp // cols, cols, S).any(1)                                        # [nQ, S]
        w = torch.zeros(nC, dtype=torch.float64, device=dev)
        for a in range(A):
            lv = torch.zeros((alive.shape[0], Sp), dtype=torch.bool, device=dev)
            lv[:, :S] = alive[:, reach[:, a]]
            bl = lv.view(-1, nC, kc).any(2).double().sum(0)                                  # [nC] live column blocks
            for o in range(O):
                zp = torch.zeros(Sp, dtype=torch.bool, device=dev)
                zp[:S] = rnz[:, a, o]
                w += zp.view(nC, kc).any(1).double() * bl
        return 2.0 * rows * kc * cols * float((G * w).sum())

    for rows, kc, cols in [(16, 16, 256), (16, 16, 128), (16, 16, 64), (16, 16, 32), (8, 16, 256), (8, 16, 64), (16, 8, 256), (16, 8, 64),
                           (8, 8, 64), (16, 4, 64), (8, 8, 32)]:
        e = executed(rows, kc, cols)
        print(f'granule {rows:2d} beliefs x {kc:2d} states x {cols:3d} alphas: {e:.4e} flops = {e / dense:.4%} of dense, {e / ideal:.2f} x element bound')


if __name__ == '__main__':
    main()
