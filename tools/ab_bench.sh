#!/bin/bash
# A/B timing of engine builds on the same box and the same saved workload: tools/ab_bench.sh lib1.so lib2.so ...
# (libraries built with PBVI_B200_LIB=... PBVI_B200_DEFS=... python -m pomdp_pbvi_exploration_b200.build --force)
set -u
python bench.py --steps 2 --warmup 1 --legs backup --no-e2e --no-cpu-baseline --save-workload /tmp/ab_wl.pt > /dev/null 2>&1
for rep in 1 2; do
  for lib in "$@"; do
    PBVI_B200_LIB=$(realpath "$lib") python bench.py --steps 20 --warmup 3 --legs backup --no-e2e --no-cpu-baseline --load-workload /tmp/ab_wl.pt 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); p=d['value_function_points']
print('$lib rep$rep ' + ' | '.join('%s step %.3f ms score %.3f ms %.2f TF (%.3g executed)' % (k, v['ms_per_step'], v['kernel_ms'], v['executed_tflops'], v['executed_flops_per_launch']) for k, v in p.items()))"
  done
done
