#!/usr/bin/env python
"""
bench.py -- belief x alpha backups/sec of the PBVI backup on the olfactory-navigation POMDP (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--beliefs B] [--alphas V]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...          (one rank per GPU)

Workload (config.workload): BASELINE.json configs[2] -- the 22021-state toroidal olfactory model (A=6, O=3, R=1),
B = 10 000 belief points from Perseus random walks (100 walks x 100 steps from b0, the engine's own expand_perseus),
V = 1 000 alpha vectors grown by the engine's own new-points backups over those walks.  One step = one full
`PBVI_Solver.backup(model, belief_set, value_function)` pass = B*V belief x alpha units: v* per (b,a,o), a*, assembly of
the distinct alpha rows and the byte-dedup that forms the new alpha set.  At N > 1 every rank backs up its own B
beliefs (weak scaling) and the ranks all-gather + merge their new alpha rows (pomdp_pbvi_exploration_b200.parallel).

`value`      device-resident inputs, CUDA-event timed, max over ranks.
`e2e`        the same call from HOST (pinned) buffers: H2D of beliefs and alphas, backup, D2H of the new alpha rows and
             actions, every step.
`roofline`   the score kernel (block-sparse FP64 DMMA GEMM + fused argmax): ALGORITHMIC flops 2*A*O*S per unit over its
             CUDA-event time, against the FP64 tensor-pipe peak measured by tools/fp64_microbench.cu on this pool
             (profiles/r01_fp64_pipe_microbench.txt; MEASURED_PEAKS.json holds no FP64 figure).
`cpu_baseline` the NumPy restatement of the reference's backup (oracle/pbvi_oracle.py, same primitives as the
             reference) on the host cores, on a bounded sample.
--impl reference runs that CPU path as the reference arm.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = 'belief x alpha backups/sec (olfactory POMDP S=22021, PBVI backup)'
UNIT = 'belief*alpha pairs/s'
GAMMA = 0.99
FP64_PEAK_TFLOPS = 37.1        # DMMA m8n8k4 / m16n8k16 on this pool's B200, profiles/r01_fp64_pipe_microbench.txt
# dram__bytes_read.sum + dram__bytes_write.sum of one score_kernel launch on the default workload (B=10000, V=1000), from
# `ncu --set full` (profiles/r01_score_kernel_v12_ncu_summary.txt); dense compulsory bytes would be 8*S*(B+V) = 1.94e9 -- chunks
# that are skipped are never read, so the kernel moves far less than that
NCU_TRAFFIC_BYTES_PER_LAUNCH = 0.2346e9


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--beliefs', type=int, default=10000)
    ap.add_argument('--alphas', type=int, default=1000)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--trace-phases', action='store_true', help='N > 1: print the phase times of the last sharded step to stderr')
    ap.add_argument('--no-dense-variant', action='store_true', help='skip the secondary measurement against a dense value function')
    ap.add_argument('--no-e2e', action='store_true', help='skip the host-buffer leg (used for short ncu passes)')
    ap.add_argument('--save-workload', default=None, help='write the synthetic beliefs / alphas to this .pt file')
    ap.add_argument('--load-workload', default=None, help='read them back instead of regenerating (ncu passes: no setup kernels)')
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """
    nvidia-smi clocks / throttle reasons, sampled every 100 ms from the warm-up on.  `begin()` / `end()` bracket the timed
    regions; the reported SM clock is the median of the samples that fall inside them (under load), or -- when a region is
    shorter than the sampling period -- of the samples of the 300 ms after it; throttle reasons come from the same samples.
    """
    FIELDS = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
              'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []          # (arrival time, line)
        self.windows = []        # [begin, end] of the timed regions

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), f'--query-gpu={self.FIELDS}', '--format=csv,noheader,nounits',
                                          '-lms', '100'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def begin(self):
        self.windows.append([time.perf_counter(), None])

    def end(self):
        if self.windows and self.windows[-1][1] is None:
            self.windows[-1][1] = time.perf_counter()

    def stop(self) -> dict:
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.3)
        self.proc.terminate()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        inside, after, mx, reasons = [], [], [], set()
        for t, ln in list(self.lines):
            parts = [p.strip() for p in ln.split(',')]
            if len(parts) < 7:
                continue
            try:
                clk, top = float(parts[0]), float(parts[1])
            except ValueError:
                continue
            mx.append(top)
            if any(b <= t <= (e if e is not None else t) for b, e in self.windows):
                inside.append(clk)
            elif any(e is not None and e < t <= e + 0.3 for b, e in self.windows):
                after.append(clk)
            else:
                continue                                   # warm-up / set-up sample: neither its clock nor its reasons are reported
            for name, val in zip(names, parts[3:7]):
                if val.lower().startswith('active'):
                    reasons.add(name)
        sm = inside if inside else after
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None, 'samples': len(sm),
                'samples_total': len(mx), 'sampled': 'inside the timed regions' if inside else 'within 300 ms after a timed region',
                'reasons': sorted(reasons)}


# ---------------------------------------------------------------------------------------------------------------------
def build_workload(model, n_beliefs: int, n_alphas: int, seed: int):
    """Synthetic inputs made by the engine itself: Perseus-walk beliefs and a value function grown by new-points backups."""
    import torch
    from pomdp_pbvi_exploration_b200 import Belief, BeliefSet, PBVI_Solver, ValueFunction
    solver = PBVI_Solver(gamma=GAMMA, eps=1e-6, expand_function='perseus')
    np.random.seed(seed)
    walks = []
    b0 = Belief(model)
    n_walks = -(-n_beliefs // 100)
    for _ in range(n_walks):
        walks.append(solver.expand_perseus(model, b0, max_generation=100).belief_array)
    beliefs = torch.cat(walks)[:n_beliefs].contiguous()
    # value function: FSVI/Perseus-style new-points backups over the walks (append=True) until V >= n_alphas
    np.random.seed(1000)                                  # the alpha set is the same on every rank
    grow_walks = [solver.expand_perseus(model, b0, max_generation=100) for _ in range(40)]
    vf = ValueFunction(model, model.expected_rewards_table.T, model.actions)
    it = 0
    while len(vf) < n_alphas and it < 400:
        vf = solver.backup(model, grow_walks[it % len(grow_walks)], vf, append=True, belief_dominance_prune=False)
        it += 1
    rows, actions = vf.alpha_vector_array, vf.actions
    if len(vf) < n_alphas:                                # top up (not expected): perturbed copies keep the shape of the data
        g = torch.Generator(device='cpu').manual_seed(7)
        need = n_alphas - len(vf)
        pick = torch.randint(0, len(vf), (need,), generator=g)
        scale = 1.0 + 1e-3 * torch.rand((need, 1), generator=g, dtype=torch.float64)
        rows = torch.cat([rows, rows[pick.to(rows.device)] * scale.to(rows.device)])
        actions = np.concatenate([actions, actions[pick.numpy()]])
    rows, actions = rows[:n_alphas].contiguous(), actions[:n_alphas]
    vf = ValueFunction(model, rows, actions)
    return solver, beliefs, vf, it


def cpu_reference_sample(model, beliefs_host: np.ndarray, alphas_host: np.ndarray, n_sample: int, n_full: int):
    """
    The reference's backup arithmetic (NumPy restatement, oracle/pbvi_oracle.py) on a bounded sample: the Gamma projection
    for ALL alphas (its cost does not depend on B) + the per-belief part for `n_sample` beliefs, extrapolated linearly in B
    to the full step (rows are independent given V).  Returns (pairs_per_s_full_step, detail dict).
    """
    from oracle import pbvi_oracle as orc
    reach = model.reachable_states
    rto = model.reachable_transitional_observation_table
    rbar = model.expected_rewards_table
    V = alphas_host.shape[0]
    A, O, S = model.action_count, model.observation_count, model.state_count
    t0 = time.perf_counter()
    G = orc.gamma_projection(reach, rto, alphas_host, GAMMA)
    t_gamma = time.perf_counter() - t0
    t0 = time.perf_counter()
    for i0 in range(0, n_sample, 64):
        b = beliefs_host[i0:min(n_sample, i0 + 64)]
        scores = np.tensordot(b, G, (1, 3))
        v_star = np.argmax(scores, axis=3)
        best_per_o = G[np.arange(A)[None, :, None, None], np.arange(O)[None, None, :, None], v_star[:, :, :, None], np.arange(S)[None, None, None, :]]
        alpha_a = rbar.T + np.sum(best_per_o, axis=2)
        values = np.einsum('bas,bs->ba', alpha_a, b)
        a_star = np.argmax(values, axis=1)
        rows = np.take_along_axis(alpha_a, a_star[:, None, None], axis=1)[:, 0, :]
        orc.dedup_rows(rows, a_star)
    t_rows = time.perf_counter() - t0
    t_full = t_gamma + t_rows * (n_full / n_sample)
    return n_full * V / t_full, {'gamma_projection_s': round(t_gamma, 3), 'per_belief_part_s': round(t_rows, 3), 'sample_beliefs': n_sample,
                                 'sample_pairs_per_s_raw': n_sample * V / (t_gamma + t_rows)}


# ---------------------------------------------------------------------------------------------------------------------
def run_reference(args):
    """Reference arm: the reference's CPU algorithm (oracle port; the Python reference itself cannot travel to the GPU box)."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    import torch
    from pomdp_pbvi_exploration_b200.recipes import olfactory_wrap_model, perseus_walk_beliefs
    model = olfactory_wrap_model()
    cores = os.cpu_count()
    torch.set_num_threads(cores)
    n_sample = 32
    beliefs = perseus_walk_beliefs(model, n_sample, seed=0)
    rng = np.random.default_rng(0)
    # alpha set of the same shape as the b200 arm's (values do not change the reference's cost): Rbar rows + smooth random rows
    alphas = np.concatenate([model.expected_rewards_table.T, rng.random((args.alphas - model.action_count, model.state_count)) * 0.1])
    vals = []
    for step in range(args.warmup + args.steps):
        v, detail = cpu_reference_sample(model, beliefs, alphas, n_sample, args.beliefs)
        if step >= args.warmup:
            vals.append((v, detail))
    value = float(np.mean([v for v, _ in vals]))
    detail = vals[-1][1]
    t_step = args.beliefs * args.alphas / value
    sample = (f'per step: Gamma projection for all {args.alphas} alphas + per-belief part on {n_sample} of {args.beliefs} beliefs, '
              f'extrapolated linearly in B; NumPy/OpenBLAS, {cores} threads')
    emit(({
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': t_step * 1e3, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': workload_config(args, 1),
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample': sample, **detail},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }))


def workload_config(args, world):
    return {'workload': f'olfactory_wrap S=22021 A=6 O=3 R=1 (BASELINE configs[2]): Perseus-walk beliefs B={args.beliefs}/GPU x V={args.alphas} alphas, '
                        f'full PBVI backup incl. dedup', 'beliefs_per_gpu': args.beliefs, 'alphas': args.alphas, 'gamma': GAMMA,
            'parallelism': f'belief-sharded x{world}' if world > 1 else 'single GPU',
            'l2_policy': 'inputs larger than L2 (beliefs 1.76 GB, alphaT 0.18 GB per step)'}


def run_b200(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    from pomdp_pbvi_exploration_b200 import BeliefSet, ValueFunction
    from pomdp_pbvi_exploration_b200.parallel import ShardedBackup
    from pomdp_pbvi_exploration_b200.recipes import olfactory_wrap_model

    model = olfactory_wrap_model()
    dev = model.device
    if args.load_workload:
        from pomdp_pbvi_exploration_b200 import PBVI_Solver
        blob = torch.load(args.load_workload)
        solver = PBVI_Solver(gamma=GAMMA, eps=1e-6, expand_function='perseus')
        beliefs = blob['beliefs'].to(dev.device)
        vf = ValueFunction(model, blob['alphas'].to(dev.device), blob['actions'].numpy())
        grow_iters = int(blob['grow_iters'])
    else:
        solver, beliefs, vf, grow_iters = build_workload(model, args.beliefs, args.alphas, seed=rank)
    if args.save_workload and rank == 0:
        torch.save({'beliefs': beliefs.cpu(), 'alphas': vf.alpha_vector_array.cpu(), 'actions': torch.as_tensor(vf.actions), 'grow_iters': grow_iters},
                   args.save_workload)
    B, V = beliefs.shape[0], len(vf)
    A, O, S = model.action_count, model.observation_count, model.state_count
    belief_set = BeliefSet(model, beliefs)
    sharded = ShardedBackup(solver, model) if world > 1 else None
    if sharded is not None:
        sharded.set_capacity(B)            # every rank backs up exactly B beliefs

    def step_device():
        if sharded is not None:
            return sharded.backup(belief_set, vf, append=False)
        return solver.backup(model, belief_set, vf, append=False, belief_dominance_prune=False)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing -----------------------------------------------------------------------------------
    sampler = ClockSampler(local_rank)
    sampler.start()                     # started before the warm-up so that nvidia-smi is up when the timed regions run
    for _ in range(args.warmup):
        out = step_device()
    dev.set_profiling(True)
    barrier()
    launches0 = dev.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    score_ms = []
    sampler.begin()
    ev0.record()
    for _ in range(args.steps):
        out = step_device()
        score_ms.append(dev.last_score_ms())       # the step has already synchronised (dedup reads keys back)
    ev1.record()
    barrier()
    sampler.end()
    if args.trace_phases and sharded is not None:
        sharded.trace = True
        step_device()
        sharded.trace = False
        print(f'[rank {rank}] phases (ms): ' + json.dumps({k: round(v, 3) for k, v in sharded.last_phases.items()}), file=sys.stderr)
    launches = dev.launch_count - launches0
    elapsed_ms = ev0.elapsed_time(ev1)
    stats = dev.last_stats()
    n_new = len(out)

    # ---- the same step against a DENSE value function (no alpha-side zeros to skip): same beliefs, same V count ------------
    dense = None
    if not args.no_dense_variant:
        g = torch.Generator(device='cpu').manual_seed(11)
        noise = (1e-3 * torch.rand(vf.alpha_vector_array.shape, generator=g, dtype=torch.float64) + 1e-6).to(dev.device)
        vf_dense = ValueFunction(model, vf.alpha_vector_array + noise, vf.actions)
        dsteps = max(2, args.steps // 2)
        for _ in range(2):
            solver.backup(model, belief_set, vf_dense, append=False, belief_dominance_prune=False)
        barrier()
        d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        dscore = []
        sampler.begin()
        d0.record()
        for _ in range(dsteps):
            solver.backup(model, belief_set, vf_dense, append=False, belief_dominance_prune=False)
            dscore.append(dev.last_score_ms())
        d1.record()
        barrier()
        sampler.end()
        dstats = dev.last_stats()
        dms = d0.elapsed_time(d1) / dsteps
        dense = {'what': 'same beliefs, the same alphas plus a strictly positive perturbation (every alpha non-zero at every state): only '
                         'belief / observation zeros are left to skip; local backup only (no exchange at N > 1)',
                 'value_per_gpu': float(B) * V / (dms * 1e-3), 'ms_per_step': dms, 'kernel_ms': float(np.mean(dscore)),
                 'executed_flops_per_launch': dstats['executed_flops'],
                 'executed_tflops': dstats['executed_flops'] / (float(np.mean(dscore)) * 1e-3) / 1e12}

    clocks = sampler.stop()       # before the e2e leg: that one is PCIe-bound, the GPU idles through most of it

    # ---- end to end from host buffers -----------------------------------------------------------------------------
    h_beliefs = beliefs.cpu().pin_memory()
    h_alphas = vf.alpha_vector_array.cpu().pin_memory()
    h_actions = vf.actions.copy()

    def step_e2e():
        bs = BeliefSet(model, h_beliefs)                             # H2D
        v_in = ValueFunction(model, h_alphas, h_actions)            # H2D (+ the constructor's byte-dedup)
        if sharded is not None:
            res = sharded.backup(bs, v_in, append=False)
        else:
            res = solver.backup(model, bs, v_in, append=False, belief_dominance_prune=False)
        rows, acts = res.numpy(staged=True)                          # D2H into the pinned staging buffer
        return rows, acts

    e2e_steps = 0 if args.no_e2e else args.steps
    for _ in range(0 if args.no_e2e else max(1, args.warmup // 2)):
        rows, acts = step_e2e()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    rows, acts = np.zeros((0, S)), np.zeros(0)
    for _ in range(e2e_steps):
        rows, acts = step_e2e()
    e1.record()
    barrier()
    e2e_ms = e0.elapsed_time(e1) if e2e_steps else 0.0   # device clock; every step ends with the blocking D2H read
    # bytes that actually crossed PCIe per step: the belief rows travel packed (bitmap + non-zero 4-double chunks, packed by host
    # threads inside the timed region and rebuilt bytewise on the device), the alpha rows as they are
    h2d_dense = int(h_beliefs.numel() * 8 + h_alphas.numel() * 8)
    h2d = int(getattr(solver, 'last_h2d_bytes', h_beliefs.numel() * 8) + h_alphas.numel() * 8) if e2e_steps else h2d_dense
    d2h = int(rows.size * 8 + acts.size * 8)

    # ---- max over ranks ---------------------------------------------------------------------------------------------
    if world > 1:
        t = torch.tensor([elapsed_ms, e2e_ms, float(np.mean(score_ms))], dtype=torch.float64, device=dev.device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms, e2e_ms, score_mean = [float(x) for x in t]
    else:
        score_mean = float(np.mean(score_ms))

    if rank == 0:
        units = float(B) * V * world
        value = units * args.steps / (elapsed_ms * 1e-3)
        e2e_value = units * args.steps / (e2e_ms * 1e-3) if e2e_ms > 0 else None
        algo_flops = 2.0 * A * O * S * B * V                  # per launch (per rank)
        achieved = algo_flops / (score_mean * 1e-3) / 1e12
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': elapsed_ms / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64',
            'data': 'synthetic', 'config': workload_config(args, world),
            'clocks': clocks,
            'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h, 'ms_per_step': (e2e_ms / args.steps) if e2e_ms > 0 else None,
                    'host_input_bytes_per_step': h2d_dense, 'host_threads': min(32, os.cpu_count() or 1) if (e2e_steps and h2d < h2d_dense) else 1,
                    'api': 'BeliefSet(host) + ValueFunction(host) -> PBVI_Solver.backup -> ValueFunction.numpy(); sparse belief rows are packed by '
                           'host threads (pbvi_pack_rows_host) for the upload and unpacked on the device'},
            'gpu_launches': int(launches),
            'roofline': {'bound': 'tensor', 'kernel': 'score_kernel<GATHER> (persistent block-sparse FP64 DMMA m8n8k4 GEMM + fused argmax)', 'achieved': achieved,
                         'peak': FP64_PEAK_TFLOPS, 'unit': 'TFLOP/s', 'frac': achieved / FP64_PEAK_TFLOPS,
                         'traffic': NCU_TRAFFIC_BYTES_PER_LAUNCH if (B == 10000 and V == 1000) else None,
                         'peak_source': 'own FP64 DMMA microbenchmark on this pool (profiles/r01_fp64_pipe_microbench.txt); '
                                        'MEASURED_PEAKS.json has no FP64 entry',
                         'algorithmic_flops_per_launch': algo_flops, 'executed_flops_per_launch': stats['executed_flops'],
                         'executed_over_algorithmic': stats['executed_flops'] / algo_flops,
                         'note': 'frac uses ALGORITHMIC (dense) flops, so it exceeds 1 by the share of exact-zero work skipped (belief, '
                                 'RTO and alpha-tile zeros; results are bit-identical to the dense computation); executed_tflops / peak '
                                 'is the pipe utilisation',
                         'executed_tflops': stats['executed_flops'] / (score_mean * 1e-3) / 1e12,
                         'kernel_ms': score_mean, 'kernel_share_of_step': score_mean / (elapsed_ms / args.steps)},
            'new_alpha_rows': n_new, 'value_function_growth_backups': grow_iters,
            'dense_alpha_variant': dense,
        }
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count()
            n_sample = 128
            v, detail = cpu_reference_sample(model, beliefs[:n_sample].cpu().numpy(), vf.alpha_vector_array.cpu().numpy(), n_sample, B)
            line['cpu_baseline'] = {'value': v, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                                    'sample': f'Gamma projection for all {V} alphas + per-belief part on {n_sample} of {B} beliefs, extrapolated '
                                              f'linearly in B; NumPy/OpenBLAS with {cores} threads', **detail}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


_JSON_OUT = None


def emit(line: dict) -> None:
    """The one JSON line goes to the real stdout; everything else printed during the run (NCCL banners, library logs) was
    redirected to stderr by main()."""
    os.write(_JSON_OUT, (json.dumps(line) + '\n').encode())


def main():
    global _JSON_OUT
    a = parse_args()
    sys.stdout.flush()
    _JSON_OUT = os.dup(1)
    os.dup2(2, 1)                      # fd 1 -> stderr for native code too (NCCL prints its version banner on stdout)
    sys.stdout = sys.stderr
    if a.impl == 'reference':
        run_reference(a)
    else:
        run_b200(a)


if __name__ == '__main__':
    main()
