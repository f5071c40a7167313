#!/bin/bash
# Dev tool: `ncu --set full` of the two kernels of the plain stepping (no energy chain) and of the heaviest kernels of the
# phase unwrapping; reports reduced to raw / details pages on the box.   usage: tools/ncu_stepping.sh <tag>
tag=${1:-final}
ncu --set full --clock-control none --kernel-name-base demangled -k "regex:col_pass_p" -s 6 -c 1 \
    -o gpurun_out/${tag}_ncu_step_col -f python tools/prof_driver.py 2048 4 > gpurun_out/${tag}_ncu_step_col.log 2>&1
ncu --set full --clock-control none --kernel-name-base demangled -k "regex:row_pass" -s 6 -c 1 \
    -o gpurun_out/${tag}_ncu_step_row -f python tools/prof_driver.py 2048 4 > gpurun_out/${tag}_ncu_step_row.log 2>&1
ncu --set full --clock-control none -k "regex:unwrap_level_union_pass" -s 0 -c 1 \
    -o gpurun_out/${tag}_ncu_unwrap_union -f python tools/unwrap_breakdown.py 2048 > gpurun_out/${tag}_ncu_unwrap_union.log 2>&1
ncu --set full --clock-control none -k "regex:unwrap_hook_pass" -s 11 -c 1 \
    -o gpurun_out/${tag}_ncu_unwrap_hook -f python tools/unwrap_breakdown.py 2048 > gpurun_out/${tag}_ncu_unwrap_hook.log 2>&1
for rep in gpurun_out/${tag}_ncu_*.ncu-rep; do
  ncu -i $rep --page raw --csv > ${rep%.ncu-rep}_raw.csv 2>/dev/null
  ncu -i $rep --page details > ${rep%.ncu-rep}_details.txt 2>/dev/null
done
rm -f gpurun_out/${tag}_ncu_*.ncu-rep
ls -la gpurun_out/${tag}_ncu_*
