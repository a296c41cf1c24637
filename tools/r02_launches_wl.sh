#!/bin/bash
# serialised launch lists of the mid-size workloads (FP64 DMMA path)
O=gpurun_out/r02w; mkdir -p $O
for wl in C2 C3; do
  ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_$wl.csv python bench.py --workload $wl --steps 2 --warmup 3 --no-predict --no-cpu-baseline > $O/ncu_$wl.log 2>&1
  python tools/summarize_launches.py $O/launches_$wl.csv > $O/launches_${wl}_summary.txt; head -16 $O/launches_${wl}_summary.txt
done
