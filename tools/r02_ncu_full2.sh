#!/bin/bash
mkdir -p gpurun_out/r02
for spec in "abar8:abar8_kernel" "kgs:kuf_grad_slice" "kuf8:kuf8_kernel"; do
  tag=${spec%%:*}; kn=${spec##*:}
  ncu --set full --import-source on --clock-control none -k regex:$kn --launch-skip 1 --launch-count 1 \
    -o gpurun_out/r02/full_$tag -f python tools/one_step.py C4 3 > gpurun_out/r02/ncu_full_$tag.log 2>&1
  ls -la gpurun_out/r02/full_$tag.ncu-rep
done
