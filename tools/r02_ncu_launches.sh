#!/bin/bash
# launch list (gpu__time_duration per launch) of one C4 step
mkdir -p gpurun_out/r02
TAG=${1:-i8step}
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02/launches_$TAG.csv \
  python bench.py --steps 2 --warmup 3 --no-predict --no-cpu-baseline > gpurun_out/r02/ncu_$TAG.log 2>&1
python tools/summarize_launches.py gpurun_out/r02/launches_$TAG.csv | tee gpurun_out/r02/launches_${TAG}_summary.txt
