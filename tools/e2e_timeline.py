"""Where the host-buffer (e2e) step of bench.py spends its time: raw link and packer rates of the box, then a per-chunk timeline of the
streamed select (packed on the host / landed on the device / selected), for several chunk plans and packer thread counts."""
import os
import subprocess
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from pomdp_pbvi_exploration_b200 import BeliefSet, ValueFunction  # noqa: E402
from pomdp_pbvi_exploration_b200.recipes import olfactory_wrap_model  # noqa: E402


def link_rates(dev, h_b):
    d = torch.empty_like(h_b, device=dev.device)
    for _ in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        d.copy_(h_b, non_blocking=True); torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    print(f'pinned H2D {h_b.numel() * 8 / 1e9:.2f} GB: {dt * 1e3:.2f} ms = {h_b.numel() * 8 / dt / 1e9:.1f} GB/s')
    back = torch.empty((1000, h_b.shape[1]), dtype=torch.float64).pin_memory()
    for _ in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        back.copy_(d[:1000], non_blocking=True); torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    print(f'pinned D2H {back.numel() * 8 / 1e9:.2f} GB: {dt * 1e3:.2f} ms = {back.numel() * 8 / dt / 1e9:.1f} GB/s')
    # both directions at once
    s2 = torch.cuda.Stream()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    d.copy_(h_b, non_blocking=True)
    with torch.cuda.stream(s2):
        for _ in range(8):
            back.copy_(d[:1000], non_blocking=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f'H2D {h_b.numel() * 8 / 1e9:.2f} GB with 8 x D2H {back.numel() * 8 / 1e9:.2f} GB alongside: {dt * 1e3:.2f} ms')
    # host memcpy rate (one thread, then torch's pool)
    src = h_b[:4000]; dst = torch.empty_like(src)
    for th in (1, 4, 8, 16):
        torch.set_num_threads(th)
        t0 = time.perf_counter(); dst.copy_(src); dt = time.perf_counter() - t0
        print(f'host copy {src.numel() * 8 / 1e9:.2f} GB with {th} torch threads: {dt * 1e3:.1f} ms = {src.numel() * 8 / dt / 1e9:.1f} GB/s read (+ same written)')


def packers_alone(dev, h_b, threads):
    """Time until the last slab is packed (nothing is copied meanwhile)."""
    type(dev).PACK_THREADS = threads
    best = 1e9
    for _ in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        job = dev.start_pack(h_b)
        for f in job.futures:
            f.result()
        best = min(best, time.perf_counter() - t0)
        job.shipped(h_b.shape[0]).synchronize()
        job.close()
    print(f'packers alone, {threads:2d} threads: {best * 1e3:7.2f} ms = {h_b.numel() * 8 / best / 1e9:.1f} GB/s scanned')


def e2e(model, solver, h_b, h_a, acts, steps=6, trace_last=True):
    times = []
    tr = None
    for it in range(steps):
        torch.cuda.synchronize()
        solver.stream_trace = tr = [] if (trace_last and it == steps - 1) else None
        e0 = torch.cuda.Event(enable_timing=True); e0.record()
        t0 = time.perf_counter()
        bs = BeliefSet(model, h_b)
        v_in = ValueFunction(model, h_a, acts)
        t1 = time.perf_counter()
        res = solver.backup(model, bs, v_in, append=False, belief_dominance_prune=False)
        t2 = time.perf_counter()
        e1 = torch.cuda.Event(enable_timing=True); e1.record()
        out = res.numpy(staged=True)
        t3 = time.perf_counter()
        times.append(((t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, (t3 - t0) * 1e3))
    solver.stream_trace = None
    t = np.array(times[1:])
    print(f'  steps (ms): {[round(x[3], 1) for x in times]}; mean of all but the first: ctors {t[:, 0].mean():.2f}  backup (host returns) {t[:, 1].mean():.2f}  '
          f'numpy(staged) {t[:, 2].mean():.2f}  step {t[:, 3].mean():.2f}')
    if tr:
        torch.cuda.synchronize()
        for c in tr:
            print(f'    rows {c["rows"][0]:5d}-{c["rows"][1]:5d}: select enqueued at {(c["enqueued_host_s"] - t0) * 1e3:6.2f} ms (host clock), landed {e0.elapsed_time(c["copied"]):6.2f} ms, '
                  f'selected {e0.elapsed_time(c["selected"]):6.2f} ms')
        print(f'    backup done on the device at {e0.elapsed_time(e1):6.2f} ms; h2d {solver.last_h2d_bytes / 1e6:.0f} MB; new rows {len(res)}')
    return float(t[:, 3].mean())


def main():
    print(subprocess.run('nproc; lscpu | grep -E "Model name|Socket|Thread|MHz|L3"; free -g | head -2', shell=True, capture_output=True, text=True).stdout)
    model = olfactory_wrap_model()
    dev = model.device
    solver, beliefs, vfs, _ = bench.build_workload(model, 10000, 1000, seed=0)
    h_b = beliefs.cpu().pin_memory()
    link_rates(dev, h_b)
    for th in (15,):
        packers_alone(dev, h_b, th)
    type(dev).PACK_THREADS = None
    for name in ('late', 'young'):
        vf = vfs[name]
        h_a = vf.alpha_vector_array.cpu().pin_memory()
        acts = vf.actions.copy()
        for early, warm in ((True, False), (True, True), (False, True), (True, True)):
            if name == 'young' and not early:
                continue
            type(solver).EARLY_ROWS = early
            print(f'{name} V, early rows {early}, GPU kept busy for 0.5 s right before the steps: {warm}:')
            if warm:
                bench.gpu_warm(dev.device, 0.5)
            else:
                time.sleep(2.0)                   # an idle GPU drops its clocks
            e2e(model, solver, h_b, h_a, acts, steps=8)
        type(solver).EARLY_ROWS = True
        type(dev).PACK_THREADS = None
        os.environ['LOCAL_WORLD_SIZE'] = '2'          # start_pack declines: the rows travel as they are
        print(f'{name} V, dense upload (no packing):')
        e2e(model, solver, h_b, h_a, acts)
        del os.environ['LOCAL_WORLD_SIZE']


if __name__ == '__main__':
    main()
