#!/bin/bash
# Dev tool: one `ncu --set full` capture per shipped hot kernel (2048^2 c128; energy chain included), reports under
# gpurun_out/<tag>_ncu_<name>.ncu-rep.     usage: tools/ncu_captures.sh <tag>
tag=${1:-final}
I='.int.'
for spec in "main_col:col_pass_p<double. ${I}2048. ${I}16. ${I}4. ${I}1. ${I}1. ${I}0. ${I}1>" \
            "main_row:row_pass<double. ${I}2048. ${I}8. ${I}1. ${I}1. ${I}1>" \
            "inv_col:col_pass_p<double. ${I}2048. ${I}16. ${I}4. ${I}1. ${I}1. ${I}0. ${I}2>" \
            "inv_row:row_pass<double. ${I}2048. ${I}8. ${I}1. ${I}1. ${I}3>" \
            "energy:energy_polar_pass"; do
  name=${spec%%:*}; rx=${spec#*:}
  ncu --set full --clock-control none --kernel-name-base demangled -k "regex:$rx" -s 2 -c 1 \
      -o gpurun_out/${tag}_ncu_${name} -f python tools/prof_energy.py > gpurun_out/${tag}_ncu_${name}.log 2>&1
done
SGPE_PREC=c64 ncu --set full --clock-control none -k "regex:col_pass_p|row_pass" -s 8 -c 2 \
      -o gpurun_out/${tag}_ncu_c64 -f python tools/prof_driver.py 2048 4 > gpurun_out/${tag}_ncu_c64.log 2>&1
# (gpurun_out/ travels back only below 64 MiB: the reports are reduced to their raw / details pages on the box)
for rep in gpurun_out/${tag}_ncu_*.ncu-rep; do
  ncu -i $rep --page raw --csv > ${rep%.ncu-rep}_raw.csv 2>/dev/null
  ncu -i $rep --page details > ${rep%.ncu-rep}_details.txt 2>/dev/null
  rm -f $rep
done
ls -la gpurun_out/${tag}_ncu_*
