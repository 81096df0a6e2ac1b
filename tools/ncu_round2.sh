#!/bin/bash
# Round-2 ncu evidence on ONE GPU (run through gpurun): launch list of a late-value-function step and `--set full` captures of the
# score kernel (late and young value functions), the assemble kernel and the belief re-layout kernel.  Numbers printed by runs under ncu
# are never bench values.  Outputs: gpurun_out/r2_launches_late.csv, gpurun_out/r2_prof_*.ncu-rep
set -u
cd "$(dirname "$0")/.."
WL=/tmp/r2_wl.pt
python bench.py --steps 2 --warmup 1 --legs backup --no-cpu-baseline --no-e2e --save-workload $WL > gpurun_out/r2_ncu_plain.json 2> gpurun_out/r2_ncu_plain.err || exit 1
CMD="python bench.py --steps 2 --warmup 1 --legs backup --no-cpu-baseline --no-e2e --load-workload $WL"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_late.csv $CMD --only-point late > gpurun_out/r2_ncu_launches.log 2>&1
echo "launch list exit $?"
ncu --set full --clock-control none --import-source on -k regex:score_kernel -s 2 -c 1 -o gpurun_out/r2_prof_score_late -f $CMD --only-point late > gpurun_out/r2_ncu_score_late.log 2>&1
echo "score late exit $?"
ncu --set full --clock-control none --import-source on -k regex:score_kernel -s 2 -c 1 -o gpurun_out/r2_prof_score_young -f $CMD --only-point young > gpurun_out/r2_ncu_score_young.log 2>&1
echo "score young exit $?"
ncu --set full --clock-control none -k regex:'assemble_grouped|belief_mask|backup_value|build_chunk_lists' -s 8 -c 4 -o gpurun_out/r2_prof_aux_late -f $CMD --only-point late > gpurun_out/r2_ncu_aux.log 2>&1
echo "aux exit $?"
ls -la gpurun_out/r2_prof_* gpurun_out/r2_launches_late.csv
