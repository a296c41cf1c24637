#!/bin/bash
# ncu --set full capture of ONE launch: $1 = kernel regex, $2 = launches to skip, $3 = tag
mkdir -p gpurun_out/r02
ncu --set full --import-source on --clock-control none -k regex:$1 --launch-skip $2 --launch-count 1 \
  -o gpurun_out/r02/full_$3 -f python tools/one_step.py C4 3 > gpurun_out/r02/ncu_full_$3.log 2>&1
tail -3 gpurun_out/r02/ncu_full_$3.log
ls -la gpurun_out/r02/full_$3.ncu-rep
