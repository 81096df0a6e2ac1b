"""python tools/ncu_summary.py report.ncu-rep [kernel-regex] -> the metrics the profiles/*_ncu_summary.txt files quote, per profiled launch."""
import csv
import io
import re
import subprocess
import sys

KEEP = re.compile(r'^(gpu__time_duration\.sum|dram__bytes_(read|write)\.sum|gpu__dram_throughput\.avg\.pct_of_peak_sustained_elapsed|'
                  r'l1tex__m_xbar2l1tex_read_bytes\.sum|launch__(block_size|grid_size|registers_per_thread)|lts__t_sector_hit_rate\.pct|'
                  r'sm__cycles_elapsed\.avg\.per_second|sm__inst_executed_pipe_tensor_subpipe_dmma\.avg\.pct_of_peak_sustained_active|'
                  r'sm__inst_executed_pipe_fp64\.avg\.pct_of_peak_sustained_active|sm__throughput\.avg\.pct_of_peak_sustained_elapsed|'
                  r'sm__warps_active\.avg\.per_cycle_active|smsp__issue_active\.avg\.pct_of_peak_sustained_active|'
                  r'smsp__warps_eligible\.avg\.per_cycle_active|smsp__inst_executed\.sum|smsp__pcsamp_warps_issue_stalled_\w+)$')


def main():
    rep = sys.argv[1]
    pat = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    k_name = hdr.index('Kernel Name')
    for r in data:
        if pat and not pat.search(r[k_name]):
            continue
        print(f'== {r[k_name]}')
        for i, h in enumerate(hdr):
            if KEEP.match(h) and r[i] not in ('', 'n/a'):
                print(f'{h:<100} {r[i]} {units[i]}')
        print()


if __name__ == '__main__':
    main()
