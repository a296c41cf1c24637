#!/bin/bash
# ncu --set full capture of one kernel (regex $1) of the C4 step -> gpurun_out/r02k/<name>.ncu-rep
O=gpurun_out/r02k
mkdir -p $O
ncu --set full --import-source on --clock-control none -k regex:"$1" --launch-skip ${3:-2} --launch-count 1 -o $O/$2 -f python tools/one_step.py C4 3 > $O/ncu_$2.log 2>&1
ls -la $O/$2.ncu-rep
