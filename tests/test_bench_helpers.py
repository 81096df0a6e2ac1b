"""CPU checks of bench.py's helpers (no GPU, no nvidia-smi needed)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


class _FakeProc:
    def terminate(self):
        pass


def _line(clk, hw='Not Active', pwr='Not Active'):
    return f'{clk}, 1965, 512.3, {hw}, Not Active, Not Active, {pwr}'


def test_clock_sampler_reports_samples_of_the_timed_regions():
    import bench
    s = bench.ClockSampler(0)
    s.proc = _FakeProc()
    t = time.perf_counter()
    s.windows = [[t + 1.0, t + 1.05], [t + 2.0, t + 2.4]]
    s.lines = [(t + 0.5, _line(300, hw='Active')),          # set-up phase: ignored, reasons included
               (t + 1.1, _line(1900)), (t + 2.1, _line(1965)), (t + 2.2, _line(1950, pwr='Active')), (t + 2.3, _line(1965)),
               (t + 3.5, _line(200)), (t + 0.6, 'not a sample')]
    out = s.stop()
    assert out['sm_mhz'] == 1965.0 and out['sm_max_mhz'] == 1965.0 and out['samples'] == 3
    assert out['reasons'] == ['sw_power_cap'] and out['sampled'] == 'inside the timed regions'
    # a region shorter than the sampling period: the samples right after it stand in
    s2 = bench.ClockSampler(0)
    s2.proc = _FakeProc()
    s2.windows = [[t + 1.0, t + 1.05]]
    s2.lines = [(t + 0.5, _line(300)), (t + 1.1, _line(1900)), (t + 1.2, _line(1965)), (t + 2.0, _line(400))]
    out2 = s2.stop()
    assert out2['samples'] == 2 and out2['sm_mhz'] == 1932.5 and out2['sampled'].startswith('within 300 ms')


def test_clock_sampler_without_nvidia_smi():
    import bench
    s = bench.ClockSampler(0)
    s.proc = None
    assert s.stop()['reasons'] == ['nvidia-smi unavailable']


def test_workload_config_names_the_baseline_config():
    import bench

    class A:
        beliefs, alphas, scaling, headline = 10000, 1000, 'weak', 'late'
    cfg = bench.workload_config(A, 8)
    assert 'BASELINE configs[2]' in cfg['workload'] and cfg['parallelism'] == 'belief-sharded x8'
    assert cfg['beliefs_per_gpu'] == 10000 and cfg['beliefs_total'] == 80000 and cfg['value_function'] == 'late'
    assert bench.workload_config(A, 1)['parallelism'] == 'single GPU'

    class Strong(A):
        beliefs, scaling = 50000, 'strong'
    cfg = bench.workload_config(Strong, 8)
    assert cfg['beliefs_per_gpu'] == 6250 and cfg['beliefs_total'] == 50000


def test_oracle_sample_and_extrapolation_on_a_small_model():
    """bench.py's oracle leg (the reference's arithmetic in 512-row chunks + the top-2 gaps of the parity contract) equals
    oracle.backup on a golden model, and the extrapolation formula is the stated one."""
    import numpy as np
    import bench
    from conftest import load_golden
    from oracle import pbvi_oracle as orc
    m, g = load_golden('model_grid4x4_noloop'), load_golden('backup_grid4x4_noloop')
    reach = m['reach'].astype(np.int64)
    out, timing = bench.oracle_backup_sample(reach, m['rto'], m['rbar'], float(m['gamma']), g['beliefs'], g['alphas'], chunk=7)
    ref = orc.backup(reach, m['rto'], m['rbar'], float(m['gamma']), g['beliefs'], g['alphas'], return_scores=True)
    assert np.array_equal(out['v_star'], ref['v_star']) and np.array_equal(out['a_star'], ref['a_star'])
    assert np.array_equal(out['alpha'], ref['alpha'])
    top = np.sort(ref['scores'], axis=3)
    assert np.allclose(out['best'], top[..., -1]) and np.allclose(out['gap'], top[..., -1] - top[..., -2])
    ext = bench.extrapolate({'gamma_projection_s': 2.0, 'per_belief_part_s': 3.0, 'sample_beliefs': 100}, 1000, 50)
    assert ext['extrapolation_factor'] == 10.0 and abs(ext['full_step_s_estimate'] - 32.0) < 1e-12
    assert abs(ext['value'] - 1000 * 50 / 32.0) < 1e-9 and abs(ext['sample_pairs_per_s_raw'] - 100 * 50 / 5.0) < 1e-9
